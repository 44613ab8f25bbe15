import os, sys, subprocess
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

def run():
    import torch, torch.nn.functional as F
    from lycoris_b200.engine import kernels as k
    torch.manual_seed(3)
    Nb,H,W,C,O,R,pad = 8,32,32,1280,1280,3,1
    dt=torch.bfloat16
    x = torch.randn(Nb,C,H,W,device='cuda',dtype=dt).contiguous(memory_format=torch.channels_last)
    w = (torch.randn(O,C,R,R,device='cuda')/(C*R*R)**0.5).to(dt)
    b = torch.zeros(O,device='cuda',dtype=dt)
    wk = w.permute(0,2,3,1).reshape(O,R*R*C)
    ref = F.conv2d(x.float(), w.float(), None, padding=pad)
    out=[]
    for rep in range(3):
        y = k.conv2d_fprop(x, wk, None, R, R, (pad,pad), 1)
        out.append(((y.float()-ref).abs().max()/ref.abs().max()).item())
    yb = k.conv2d_fprop(x, wk, b, R, R, (pad,pad), 1)
    out.append(('bias0', ((yb.float()-ref).abs().max()/ref.abs().max()).item()))
    # second data set (like dgrad): different tensors
    dy = torch.randn(Nb,O,H,W,device='cuda',dtype=dt).contiguous(memory_format=torch.channels_last)
    wd = w.flip(2,3).permute(1,2,3,0).reshape(C,R*R*O)
    refd = F.conv2d(dy.float(), w.flip(2,3).permute(1,0,2,3).float(), None, padding=pad)
    for rep in range(3):
        dx = k.conv2d_fprop(dy, wd, None, R, R, (pad,pad), 1)
        torch.cuda.synchronize()
        e = (dx.float()-refd).abs()
        bad = (e > 0.05*refd.abs().max())
        cols = bad.any(dim=0).any(dim=-1).any(dim=-1).nonzero().flatten()
        out.append(('dgrad', (e.max()/refd.abs().max()).item(), int(bad.sum()), cols[:6].tolist(), cols[-3:].tolist()))
    print(os.environ.get('LYCO_CONV_BN'), out, flush=True)

if __name__ == '__main__':
    if len(sys.argv) > 1:
        run()
    else:
        for bn in ('256','224','192','128','64'):
            env = dict(os.environ, LYCO_CONV_BN=bn)
            subprocess.run([sys.executable, __file__, 'x'], env=env)
