"""Standalone target for `ncu --set full`: the HBM-bound layout kernels around the implicit-GEMM convolutions
(NCHW->NHWC transpose+cast, filter re-layouts) and the convolution with its three epilogues, on SDXL shapes."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from lycoris_b200.engine import kernels as k


def main():
    torch.manual_seed(0)
    reps = int(os.environ.get("REPS", "2"))
    # ResNet conv1 of the 320-channel level: fp32 NCHW activations [8, 320, 128, 128] under autocast
    x32 = torch.randn(8, 320, 128, 128, device="cuda")
    x16 = x32.to(torch.bfloat16)
    w = (torch.randn(1280, 1280, 3, 3, device="cuda") * 0.01).to(torch.bfloat16)
    dwk = torch.randn(1280, 9 * 1280, device="cuda")
    xs = torch.randn(8, 1280, 32, 32, device="cuda", dtype=torch.bfloat16).contiguous(memory_format=torch.channels_last)
    for _ in range(reps):
        k.as_nhwc(x32, torch.bfloat16)      # 168 MB read + 84 MB written
        k.as_nhwc(x16, torch.bfloat16)      # 84 + 84 MB
        wk = k.filter_relayout(w, k.FILTER_FPROP)   # 29.5 + 29.5 MB
        wd = k.filter_relayout(w, k.FILTER_DGRAD)
        k.filter_relayout((dwk, (1280, 1280, 3, 3)), k.FILTER_WBACK)   # 59 + 59 MB
        k.conv2d_fprop(xs, wk, None, 3, 3, (1, 1), 1)                                       # NHWC epilogue
        k.conv2d_fprop(xs, wk, None, 3, 3, (1, 1), 1, out_nchw=True)                        # NCHW, TMA stores
        k.conv2d_fprop(xs, wd, None, 3, 3, (1, 1), 1, out_nchw=True, out_dtype=torch.float32)  # NCHW fp32
    # round 2: DoRA rescale around the merged weight and the standalone delta weight, at the GEGLU projection's size
    Wm = (torch.randn(10240, 1280, device="cuda") * 0.03).to(torch.bfloat16)
    g = torch.rand(10240, device="cuda") + 0.5
    dW = torch.randn(10240, 1280, device="cuda")
    w1 = torch.randn(8, 8, device="cuda") * 0.3
    w2 = torch.randn(1280, 160, device="cuda") * 0.02
    d = k.make_desc(k.ALGO_LOKR, 10240, 1280, factors=[w1, w2], w_dtype=torch.float32, up=8, uq=8, vp=1280, vq=160)
    for _ in range(reps):
        out, sumsq = k.dora_fwd(Wm, g, True, 1, 1.0, 1.19e-7)
        k.dora_bwd(dW, Wm, g, sumsq, True, 1, 1.0, 1.19e-7)
        k.delta_weight(d, (10240, 1280), torch.float32, None, True, True)
    torch.cuda.synchronize()
    print("ncu layout target done")


if __name__ == "__main__":
    main()
