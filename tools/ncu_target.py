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
    for (M, N, K) in ((8192, 10240, 1280), (8192, 1280, 1280)):
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
        torch.cuda.synchronize()
    print("ncu target done")


if __name__ == "__main__":
    main()
