"""Small standalone target for `ncu --set full`: the three contractions of the dominant SDXL LoKr
layer shapes plus the merge / factor-gradient kernels, a few launches each (no model, no bench)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from lycoris_b200.engine import kernels as k


def main():
    torch.manual_seed(0)
    reps = int(os.environ.get("REPS", "3"))
    for (M, N, K) in ((8192, 10240, 1280), (8192, 1280, 1280), (8192, 1280, 5120)):
        X = torch.randn(M, K, device="cuda", dtype=torch.bfloat16)
        W = (torch.randn(N, K, device="cuda") * 0.03).to(torch.bfloat16)
        b = torch.zeros(N, device="cuda", dtype=torch.bfloat16)
        dY = torch.randn(M, N, device="cuda", dtype=torch.bfloat16)
        w1 = torch.randn(8, 8, device="cuda") * 0.3
        w2 = torch.randn(N // 8, K // 8, device="cuda") * 0.02
        d = k.make_desc(k.ALGO_LOKR, N, K, factors=[w1, w2], w_dtype=torch.bfloat16, up=8, uq=8, vp=N // 8, vq=K // 8)
        for _ in range(reps):
            Wm = k.merge_weight(d, W)
            y = k.gemm(X, Wm, bias=b)
            dx = k.gemm(dY, Wm, b_mn=True)
            dw = k.gemm(dY, X, a_mn=True, b_mn=True, out_dtype=torch.float32)
            g = k.factor_grads(d, dw, None, [w1.shape, w2.shape])
            # round 2: the structured factor gradients that replace `dw` + factor_grads for Linear LoKr layers
            up = uq = 8
            vp, vq = N // 8, K // 8
            w2c = w2.to(torch.bfloat16)
            if up * vq <= uq * vp:
                Xt = k.lokr_mix(X, w1, up, uq, vq, False)
                dY2 = dY.view(M * up, vp)
                g2 = k.gemm(dY2, Xt.view(M * up, vq), a_mn=True, b_mn=True, out_dtype=torch.float32)
                Q = k.gemm(dY2, w2c.t().contiguous())
                g1 = k.lokr_w1grad(Q.view(M, up * vq), X, up, uq, vq, 1.0)
            else:
                Z = k.lokr_mix(dY, w1, uq, up, vp, True)
                X2 = X.view(M * uq, vq)
                g2 = k.gemm(Z.view(M * uq, vp), X2, a_mn=True, b_mn=True, out_dtype=torch.float32)
                H = k.gemm(X2, w2c)
                g1 = k.lokr_w1grad(dY, H.view(M, uq * vp), up, uq, vp, 1.0)
        torch.cuda.synchronize()
    # 3x3 convolution, SDXL 1280-channel block at 32x32 (batch 8): fprop, dgrad (fprop on dY), wgrad
    Nb, C, O, H = 8, 1280, 1280, 32
    x = torch.randn(Nb, C, H, H, device="cuda", dtype=torch.bfloat16).contiguous(memory_format=torch.channels_last)
    dy = torch.randn(Nb, O, H, H, device="cuda", dtype=torch.bfloat16).contiguous(memory_format=torch.channels_last)
    w = (torch.randn(O, C, 3, 3, device="cuda") * 0.01).to(torch.bfloat16)
    wk = k.filter_relayout(w, k.FILTER_FPROP)
    wd = k.filter_relayout(w, k.FILTER_DGRAD)
    for _ in range(reps):
        k.conv2d_fprop(x, wk, None, 3, 3, (1, 1), 1)
        k.conv2d_fprop(dy, wd, None, 3, 3, (1, 1), 1)
        k.conv2d_wgrad(x, dy, 3, 3, (1, 1), 1)
    torch.cuda.synchronize()
    # LoHa tile kernel (lyco_hada): forward merge and the gradient operands, rank 32, on a small and a large layer
    for (N, Kd) in ((1280, 1280), (10240, 1280)):
        r = 32
        f = [(torch.randn(N, r, device="cuda") * 0.1).to(torch.bfloat16), (torch.randn(r, Kd, device="cuda") * 0.1).to(torch.bfloat16),
             (torch.randn(N, r, device="cuda") * 0.1).to(torch.bfloat16), (torch.randn(r, Kd, device="cuda") * 0.1).to(torch.bfloat16)]
        W = (torch.randn(N, Kd, device="cuda") * 0.03).to(torch.bfloat16)
        dW = torch.randn(N, Kd, device="cuda")
        for _ in range(reps):
            k.hada_merge(f, W, 1.0, 1.0, 1.0)
            k.hada_grad_operands(f, dW, 1.0)
    torch.cuda.synchronize()
    print("ncu target done")


if __name__ == "__main__":
    main()
