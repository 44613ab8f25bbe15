"""First-contact probe of every kernel on a real B200: each case runs in its own subprocess so a
trap / illegal instruction in one kernel cannot poison the others.  Writes gpurun_out/probe.log."""
import json
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def case_gemm(M, N, K, a_mn, b_mn, out, bias, split_k=0, dtype="bf16", force=None):
    if force:
        os.environ["LYCO_GEMM_FORCE"] = force
    import torch
    from lycoris_b200.engine import kernels as k

    dt = torch.bfloat16 if dtype == "bf16" else torch.float16
    torch.manual_seed(0)
    A = torch.randn(M, K, device="cuda", dtype=dt)
    B = torch.randn(N, K, device="cuda", dtype=dt)
    a = A.t().contiguous() if a_mn else A
    b = B.t().contiguous() if b_mn else B
    bi = torch.randn(N, device="cuda", dtype=dt) if bias else None
    odt = torch.float32 if out == "f32" else dt
    ref = A.float() @ B.float().t()
    if bi is not None:
        ref = ref + bi.float()
    C = k.gemm(a, b, a_mn=a_mn, b_mn=b_mn, bias=bi, out_dtype=odt, split_k=split_k)
    torch.cuda.synchronize()
    err = (C.float() - ref).abs().max().item()
    rel = err / ref.abs().max().item()
    # timing
    for _ in range(3):
        k.gemm(a, b, a_mn=a_mn, b_mn=b_mn, bias=bi, out_dtype=odt, split_k=split_k)
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    e0.record()
    it = 10
    for _ in range(it):
        k.gemm(a, b, a_mn=a_mn, b_mn=b_mn, bias=bi, out_dtype=odt, split_k=split_k)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / it
    tf = 2.0 * M * N * K / ms / 1e9
    # cuBLAS for scale
    for _ in range(3):
        torch.matmul(A, B.t())
    e0.record()
    for _ in range(it):
        torch.matmul(A, B.t())
    e1.record()
    torch.cuda.synchronize()
    ms_ref = e0.elapsed_time(e1) / it
    return {"max_abs_err": err, "rel_err": rel, "ms": ms, "tflops": tf, "cublas_ms": ms_ref,
            "cublas_tflops": 2.0 * M * N * K / ms_ref / 1e9, "ok": bool(rel < 2e-2)}


def case_weight(algo, N, K, r, fdt="f32"):
    import torch
    from lycoris_b200.engine import kernels as k

    torch.manual_seed(1)
    dev = "cuda"
    fd = torch.float32 if fdt == "f32" else torch.bfloat16
    W = (torch.rand(N, K, device=dev) * 0.25 - 0.125).to(torch.bfloat16)
    pre_round = 1
    res = {}
    if algo in ("locon", "loha"):
        f = [torch.randn(N, r, device=dev, dtype=fd) * 0.1, torch.randn(r, K, device=dev, dtype=fd) * 0.1]
        if algo == "loha":
            f += [torch.randn(N, r, device=dev, dtype=fd) * 0.1, torch.randn(r, K, device=dev, dtype=fd) * 0.1]
        scale = 0.5
        desc = k.make_desc(k.ALGO_LOCON if algo == "locon" else k.ALGO_LOHA, N, K, factors=f,
                           w_dtype=torch.bfloat16, rank=r, pre_round=1, pre_dtype=torch.bfloat16,
                           m_pre=(1.0 if algo == "locon" else scale), m_post1=(scale if algo == "locon" else 1.0))
        fb = [t.to(torch.bfloat16).float() for t in f]
        if algo == "locon":
            raw = (fb[0] @ fb[1]).to(torch.bfloat16)
            d = (raw * scale)
        else:
            raw = ((fb[0] @ fb[1]).to(torch.bfloat16) * (fb[2] @ fb[3]).to(torch.bfloat16))
            d = (raw * scale)
        ref = (W + d)
        shapes = [t.shape for t in f]
    elif algo == "lokr":
        up, uq = 8, 8
        vp, vq = N // up, K // uq
        f = [torch.randn(up, uq, device=dev, dtype=fd) * 0.3, torch.randn(vp, vq, device=dev, dtype=fd) * 0.05]
        desc = k.make_desc(k.ALGO_LOKR, N, K, factors=f, w_dtype=torch.bfloat16, up=up, uq=uq, vp=vp, vq=vq,
                           pre_round=0)
        d = torch.kron(f[0].float(), f[1].float()).to(torch.bfloat16)
        ref = W + d
        shapes = [t.shape for t in f]
    elif algo == "ia3":
        f = [torch.randn(N, device=dev, dtype=fd) * 0.1]
        desc = k.make_desc(k.ALGO_IA3, N, K, factors=f, w_dtype=torch.bfloat16, on_input=0)
        ref = (W.float() * (f[0].float() + 1)[:, None]).to(torch.bfloat16)
        shapes = [t.shape for t in f]
    out = k.merge_weight(desc, W)
    torch.cuda.synchronize()
    res["merge_max_err"] = (out.float() - ref.float()).abs().max().item()
    res["merge_mismatch_frac"] = (out != ref).float().mean().item()
    # grads vs autograd in fp32
    dW = torch.randn(N, K, device=dev, dtype=torch.float32)
    fr = [t.detach().clone().float().requires_grad_(True) for t in f]
    if algo == "locon":
        dd = (fr[0].to(torch.bfloat16).float() @ fr[1].to(torch.bfloat16).float()) * 0.5
    elif algo == "loha":
        dd = (fr[0] @ fr[1]) * (fr[2] @ fr[3]) * 0.5
    elif algo == "lokr":
        dd = torch.kron(fr[0], fr[1])
    else:
        dd = W.float() * (fr[0] + 1)[:, None]
    (dd * dW).sum().backward()
    gs = k.factor_grads(desc, dW, W, shapes)
    torch.cuda.synchronize()
    errs = []
    for g, t in zip(gs, fr):
        ref_g = t.grad
        errs.append(((g - ref_g).abs().max() / (ref_g.abs().max() + 1e-12)).item())
    res["grad_rel_errs"] = errs
    # timing
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    for fn, name, nbytes in ((lambda: k.merge_weight(desc, W), "merge", 4.0 * N * K),
                             (lambda: k.factor_grads(desc, dW, W, shapes), "grad", 4.0 * N * K)):
        for _ in range(3):
            fn()
        e0.record()
        for _ in range(10):
            fn()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 10
        res[name + "_ms"] = ms
        res[name + "_GBps"] = nbytes / ms / 1e6
    res["ok"] = bool(res["merge_max_err"] < 2e-2 and max(errs) < 3e-2)
    return res


def case_conv(Nb, H, W, C, O, R, pad, stride):
    import torch
    import torch.nn.functional as F
    from lycoris_b200.engine import kernels as k

    torch.manual_seed(3)
    dt = torch.bfloat16
    x = torch.randn(Nb, C, H, W, device="cuda", dtype=dt).contiguous(memory_format=torch.channels_last)
    w = (torch.randn(O, C, R, R, device="cuda") / (C * R * R) ** 0.5).to(dt)
    b = torch.randn(O, device="cuda", dtype=dt)
    res = {}
    wk = w.permute(0, 2, 3, 1).reshape(O, R * R * C)
    y = k.conv2d_fprop(x, wk, b, R, R, (pad, pad), stride)
    torch.cuda.synchronize()
    xr = x.float().requires_grad_(True)
    wr = w.float().requires_grad_(True)
    ref = F.conv2d(xr, wr, b.float(), stride=stride, padding=pad)
    res["fprop_rel"] = ((y.float() - ref).abs().max() / ref.abs().max()).item()
    dy = torch.randn_like(ref).to(dt).contiguous(memory_format=torch.channels_last)
    ref.backward(dy.float())
    dwk = k.conv2d_wgrad(x, dy, R, R, (pad, pad), stride)
    dw = dwk.view(O, R, R, C).permute(0, 3, 1, 2)
    res["wgrad_rel"] = ((dw - wr.grad).abs().max() / wr.grad.abs().max()).item()
    ok = res["fprop_rel"] < 1e-2 and res["wgrad_rel"] < 1e-3
    if stride == 1 and O % 64 == 0:
        wd = w.flip(2, 3).permute(1, 2, 3, 0).reshape(C, R * R * O)
        dx = k.conv2d_fprop(dy, wd, None, R, R, (R - 1 - pad, R - 1 - pad), 1)
        res["dgrad_rel"] = ((dx.float() - xr.grad).abs().max() / xr.grad.abs().max()).item()
        ok = ok and res["dgrad_rel"] < 1e-2
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    flops = 2.0 * y.numel() * C * R * R
    for name, fn in (("fprop", lambda: k.conv2d_fprop(x, wk, b, R, R, (pad, pad), stride)),
                     ("wgrad", lambda: k.conv2d_wgrad(x, dy, R, R, (pad, pad), stride)),
                     ("cudnn_fprop", lambda: F.conv2d(x, w, b, stride=stride, padding=pad)),
                     ("cudnn_bwd", lambda: torch.ops.aten.convolution_backward(dy, x, w, None, [stride] * 2, [pad] * 2, [1, 1], False, [0, 0], 1, [True, True, False]))):
        for _ in range(3):
            fn()
        e0.record()
        for _ in range(5):
            fn()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 5
        res[name + "_ms"] = ms
        res[name + "_tflops"] = flops * (2 if name == "cudnn_bwd" else 1) / ms / 1e9
    res["ok"] = bool(ok)
    return res


CASES = []
for bn_shape in [(256, 256, 128), (128, 64, 64), (1000, 328, 200)]:
    CASES.append(("gemm", dict(M=bn_shape[0], N=bn_shape[1], K=bn_shape[2], a_mn=False, b_mn=False, out="16", bias=False)))
CASES += [
    ("gemm", dict(M=8192, N=1280, K=1280, a_mn=False, b_mn=False, out="16", bias=True)),
    ("gemm", dict(M=512, N=512, K=256, a_mn=False, b_mn=True, out="16", bias=False)),
    ("gemm", dict(M=8192, N=1280, K=1280, a_mn=False, b_mn=True, out="16", bias=False)),
    ("gemm", dict(M=512, N=512, K=256, a_mn=True, b_mn=True, out="f32", bias=False, split_k=1)),
    ("gemm", dict(M=1280, N=1280, K=8192, a_mn=True, b_mn=True, out="f32", bias=False, split_k=1)),
    ("gemm", dict(M=1280, N=1280, K=8192, a_mn=True, b_mn=True, out="f32", bias=False, split_k=0)),
    ("gemm", dict(M=512, N=512, K=256, a_mn=True, b_mn=False, out="16", bias=False)),
    ("gemm", dict(M=8192, N=10240, K=1280, a_mn=False, b_mn=False, out="16", bias=True)),
    ("gemm", dict(M=32768, N=640, K=640, a_mn=False, b_mn=False, out="16", bias=True)),
    ("gemm", dict(M=616, N=1280, K=2048, a_mn=False, b_mn=False, out="16", bias=False)),
    ("gemm", dict(M=1024, N=512, K=512, a_mn=False, b_mn=False, out="16", bias=True, dtype="f16")),
    # CTA-pair (cta_group::2) variants
    ("gemm", dict(M=512, N=512, K=256, a_mn=False, b_mn=False, out="16", bias=True, force="pair256")),
    ("gemm", dict(M=512, N=512, K=256, a_mn=False, b_mn=False, out="16", bias=True, force="pair128")),
    ("gemm", dict(M=1000, N=328, K=200, a_mn=False, b_mn=False, out="16", bias=False, force="pair256")),
    ("gemm", dict(M=512, N=512, K=256, a_mn=False, b_mn=True, out="16", bias=False, force="pair256")),
    ("gemm", dict(M=512, N=512, K=256, a_mn=True, b_mn=True, out="f32", bias=False, split_k=1, force="pair256")),
    ("gemm", dict(M=512, N=512, K=256, a_mn=True, b_mn=True, out="f32", bias=False, split_k=3, force="pair128")),
    ("gemm", dict(M=8192, N=1280, K=1280, a_mn=False, b_mn=False, out="16", bias=True, force="pair256")),
    ("gemm", dict(M=8192, N=1280, K=1280, a_mn=False, b_mn=False, out="16", bias=True, force="pair128")),
    ("gemm", dict(M=8192, N=10240, K=1280, a_mn=False, b_mn=False, out="16", bias=True, force="pair256")),
    ("gemm", dict(M=8192, N=1280, K=5120, a_mn=False, b_mn=False, out="16", bias=True, force="pair256")),
    ("gemm", dict(M=8192, N=1280, K=1280, a_mn=False, b_mn=True, out="16", bias=False, force="pair256")),
    ("gemm", dict(M=1280, N=1280, K=8192, a_mn=True, b_mn=True, out="f32", bias=False, split_k=0, force="pair256")),
    ("gemm", dict(M=32768, N=640, K=640, a_mn=False, b_mn=False, out="16", bias=True, force="pair256")),
    ("gemm", dict(M=32768, N=640, K=640, a_mn=False, b_mn=False, out="16", bias=True, force="pair128")),
    ("gemm", dict(M=8192, N=10240, K=1280, a_mn=False, b_mn=False, out="16", bias=True)),
    ("gemm", dict(M=8192, N=1280, K=1280, a_mn=False, b_mn=False, out="16", bias=True)),
    # runtime BLOCK_N (multiples of 32) + TMA-store epilogue
    ("gemm", dict(M=1000, N=328, K=200, a_mn=False, b_mn=False, out="16", bias=True, force="pair160")),
    ("gemm", dict(M=1000, N=328, K=200, a_mn=False, b_mn=False, out="16", bias=True, force="single96")),
    ("gemm", dict(M=300, N=72, K=136, a_mn=False, b_mn=False, out="16", bias=True, force="single32")),
    ("gemm", dict(M=520, N=200, K=64, a_mn=False, b_mn=False, out="16", bias=False, force="pair64")),
    ("gemm", dict(M=8192, N=1280, K=1280, a_mn=False, b_mn=False, out="16", bias=True, force="pair160")),
    ("gemm", dict(M=8192, N=1280, K=1280, a_mn=False, b_mn=False, out="16", bias=True, force="pair224")),
    ("gemm", dict(M=8192, N=1280, K=5120, a_mn=False, b_mn=False, out="16", bias=True)),
    ("gemm", dict(M=32768, N=640, K=640, a_mn=False, b_mn=False, out="16", bias=True)),
    ("gemm", dict(M=32768, N=5120, K=640, a_mn=False, b_mn=False, out="16", bias=True)),
    ("gemm", dict(M=8192, N=1280, K=1280, a_mn=False, b_mn=True, out="16", bias=False)),
    ("gemm", dict(M=1280, N=1280, K=8192, a_mn=True, b_mn=True, out="f32", bias=False, split_k=0)),
    ("gemm", dict(M=616, N=1280, K=2048, a_mn=False, b_mn=False, out="16", bias=False)),
    ("weight", dict(algo="locon", N=1280, K=1280, r=16)),
    ("weight", dict(algo="locon", N=320, K=2880, r=8, fdt="bf16")),
    ("conv", dict(Nb=2, H=8, W=8, C=64, O=64, R=3, pad=1, stride=1)),
    ("conv", dict(Nb=2, H=12, W=10, C=128, O=192, R=3, pad=1, stride=1)),
    ("conv", dict(Nb=3, H=9, W=7, C=64, O=72, R=3, pad=1, stride=1)),
    ("conv", dict(Nb=2, H=16, W=16, C=64, O=64, R=3, pad=1, stride=2)),
    ("conv", dict(Nb=2, H=16, W=16, C=320, O=320, R=3, pad=1, stride=1)),
    ("conv", dict(Nb=8, H=32, W=32, C=1280, O=1280, R=3, pad=1, stride=1)),
    ("conv", dict(Nb=8, H=128, W=128, C=320, O=320, R=3, pad=1, stride=1)),
    ("conv", dict(Nb=8, H=64, W=64, C=640, O=640, R=3, pad=1, stride=1)),
    ("conv", dict(Nb=8, H=32, W=32, C=2560, O=1280, R=3, pad=1, stride=1)),
    ("weight", dict(algo="loha", N=1280, K=1280, r=32)),
    ("weight", dict(algo="lokr", N=1280, K=1280, r=0)),
    ("weight", dict(algo="lokr", N=10240, K=1280, r=0, fdt="bf16")),
    ("weight", dict(algo="ia3", N=1280, K=2048, r=0)),
]


def main():
    if len(sys.argv) > 1 and sys.argv[1] == "--one":
        kind, kw = CASES[int(sys.argv[2])]
        fn = {"gemm": case_gemm, "weight": case_weight, "conv": case_conv}[kind]
        print("RESULT " + json.dumps(fn(**kw)))
        return
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    log = open(os.path.join(ROOT, "gpurun_out", "probe.log"), "w")
    sel = [int(x) for x in sys.argv[1:]] or range(len(CASES))
    nfail = 0
    for i in sel:
        kind, kw = CASES[i]
        t0 = time.time()
        try:
            p = subprocess.run([sys.executable, __file__, "--one", str(i)], capture_output=True, text=True, timeout=180)
            out = p.stdout + p.stderr
            rc = p.returncode
        except subprocess.TimeoutExpired as e:
            out, rc = f"TIMEOUT {e}", -9
        res = [l for l in out.splitlines() if l.startswith("RESULT ")]
        line = f"[{i}] {kind} {kw} rc={rc} t={time.time()-t0:.1f}s " + (res[0] if res else "NO RESULT\n" + out[-1500:])
        if rc != 0 or not res or not json.loads(res[0][7:]).get("ok"):
            nfail += 1
            line = "FAIL " + line
        print(line, flush=True)
        log.write(line + "\n")
        log.flush()
    print(f"probe done: {nfail} failing of {len(list(sel))}")


if __name__ == "__main__":
    main()
