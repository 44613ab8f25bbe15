"""Timeline of ONE lyco_gemm launch: where a short GEMM's time goes besides the MMA loop.

Builds a DEBUG copy of the C-ABI library with -DLYCO_GEMM_TRACE (tools/_trace/lyco_trace.so, git-ignored; the product
library has the probes compiled out), launches the requested problems through it with a per-CTA clock64 buffer and
prints, per problem: kernel time by CUDA events (median of back-to-back launches, product-style), and from the traced
launch the median / max over CTAs of  setup, first operand latency, per-tile accumulator-ready spacing, last epilogue,
teardown — in SM cycles and converted with the cycles-per-microsecond of the launch itself (globaltimer).

    python tools/gemm_trace.py --build            # here (nvcc cross-compiles)
    gpurun -- python tools/gemm_trace.py          # on the B200
"""
import argparse
import ctypes
import os
import statistics
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
OUT_DIR = os.path.join(ROOT, "tools", "_trace")
SO = os.path.join(OUT_DIR, "lyco_trace.so")

PROBLEMS = [
    # name, M, N, K, a_mn, b_mn, c_f32
    ("fwd   8192x1280x1280", 8192, 1280, 1280, 0, 0, 0),
    ("dgrad 8192x1280x1280", 8192, 1280, 1280, 0, 1, 0),
    ("wgrad 1280x1280x8192", 1280, 1280, 8192, 1, 1, 1),
    ("fwd   32768x640x640", 32768, 640, 640, 0, 0, 0),
    ("fwd   8192x5120x1280", 8192, 5120, 1280, 0, 0, 0),
]


def build():
    import __graft_entry__ as ge

    os.makedirs(OUT_DIR, exist_ok=True)
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    cmd = [nvcc, *ge.NVCC_FLAGS, "-DLYCO_GEMM_TRACE", "-o", SO, os.path.join(ge.CSRC, "lyco_abi.cu")]
    print(" ".join(cmd), flush=True)
    subprocess.run(cmd, check=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--build", action="store_true")
    ap.add_argument("--force", default=None, help="value for LYCO_GEMM_FORCE (e.g. pair192)")
    ap.add_argument("--only", default=None, help="substring filter on the problem names")
    args = ap.parse_args()
    if args.build:
        build()
        return
    if args.force:
        os.environ["LYCO_GEMM_FORCE"] = args.force
    import torch

    lib = ctypes.CDLL(SO)
    lib.lyco_gemm.restype = ctypes.c_int
    lib.lyco_last_error.restype = ctypes.c_char_p
    P, I, L = ctypes.c_void_p, ctypes.c_int, ctypes.c_int64
    lib.lyco_gemm.argtypes = [P, I, L, P, I, L, P, I, L, P, I, I, I, I, I, I, I, P]
    lib.lyco_debug_set_trace.argtypes = [P]
    dev = torch.device("cuda:0")
    stream = torch.cuda.current_stream().cuda_stream

    for name, M, N, K, a_mn, b_mn, c_f32 in PROBLEMS:
        if args.only and args.only not in name:
            continue
        a = torch.randn((K, M) if a_mn else (M, K), device=dev).to(torch.bfloat16)
        b = torch.randn((K, N) if b_mn else (N, K), device=dev).to(torch.bfloat16)
        c = torch.empty((M, N), device=dev, dtype=torch.float32 if c_f32 else torch.bfloat16)

        def launch():
            rc = lib.lyco_gemm(a.data_ptr(), a_mn, a.stride(0), b.data_ptr(), b_mn, b.stride(0), c.data_ptr(),
                               2 if c_f32 else 0, c.stride(0), None, 0, M, N, K, 0, 0, 0, stream)
            if rc:
                raise RuntimeError(lib.lyco_last_error().decode())

        lib.lyco_debug_set_trace(None)
        for _ in range(5):
            launch()
        torch.cuda.synchronize()
        times = []
        for _ in range(7):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(20):
                launch()
            e1.record()
            torch.cuda.synchronize()
            times.append(e0.elapsed_time(e1) / 20 * 1e3)
        us = statistics.median(times)
        ref = None
        if not c_f32:
            at = a.t() if a_mn else a
            bt = b if b_mn else b.t()
            for _ in range(3):
                torch.matmul(at, bt)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(20):
                torch.matmul(at, bt)
            e1.record()
            torch.cuda.synchronize()
            ref = e0.elapsed_time(e1) / 20 * 1e3

        trace = torch.zeros((512, 64), device=dev, dtype=torch.int64)
        lib.lyco_debug_set_trace(trace.data_ptr())
        launch()
        torch.cuda.synchronize()
        lib.lyco_debug_set_trace(None)
        t = trace.cpu()
        rows = [t[i] for i in range(512) if int(t[i][0]) != 0]
        cyc_per_us = statistics.median(
            [(int(r[40]) - int(r[0])) / max(1e-3, (int(r[41]) - int(r[1])) / 1e3) for r in rows if int(r[41]) > int(r[1])] or [1900.0])
        gt0 = min(int(r[1]) for r in rows)
        leaders = [r for r in rows if int(r[6]) != 0]  # CTAs that issued MMAs
        epi = [r for r in rows if int(r[42]) > 0]

        def stat(vals):
            vals = [v for v in vals if v is not None]
            if not vals:
                return "      n/a"
            return f"med {statistics.median(vals) / cyc_per_us:6.2f} max {max(vals) / cyc_per_us:6.2f} us"

        tf = 2.0 * M * N * K / us / 1e6
        print(f"\n== {name}  a_mn={a_mn} b_mn={b_mn} f32={c_f32}: {us:.2f} us/launch = {tf:.0f} TFLOP/s"
              + (f"   (torch.matmul {ref:.2f} us = {2.0 * M * N * K / ref / 1e6:.0f})" if ref else ""))
        print(f"   CTAs {len(rows)}  cycles/us {cyc_per_us:.0f}  start skew (globaltimer) "
              f"{max(int(r[1]) for r in rows) - gt0} ns  traced launch wall {max(int(r[41]) for r in rows) - gt0} ns")
        print("   entry -> setup done          ", stat([int(r[2]) - int(r[0]) for r in rows]))
        print("   setup done -> first operands ", stat([int(r[4]) - int(r[2]) for r in leaders]))
        print("   entry -> last TMA issued     ", stat([int(r[5]) - int(r[0]) for r in rows if int(r[5])]))
        print("   entry -> last MMA commit     ", stat([int(r[6]) - int(r[0]) for r in leaders]))
        ntile = max(int(r[42]) for r in epi)
        for i in range(min(ntile, 8)):
            have = [r for r in epi if int(r[42]) > i]
            print(f"   tile {i}: entry -> acc ready   ", stat([int(r[8 + 2 * i]) - int(r[0]) for r in have]),
                  " drain", stat([int(r[9 + 2 * i]) - int(r[8 + 2 * i]) for r in have]), f" ({len(have)} CTAs)")
        print("   last drain end -> exit       ",
              stat([int(r[40]) - int(r[9 + 2 * (min(int(r[42]), 16) - 1)]) for r in epi]))
        print("   entry -> exit                ", stat([int(r[40]) - int(r[0]) for r in rows]))
        del a, b, c


if __name__ == "__main__":
    main()
