#!/usr/bin/env python
"""bench.py — SDXL-UNet + LoKr (factor 8, full-dim) fwd+bwd step rate on N x B200.

    python bench.py --gpus N --steps K --warmup W          # this engine (default arm)
    python bench.py --impl reference ...                   # the reference's CPU path (oracle port)
    python bench.py --impl reference-gpu ...               # informational: reference-equivalent eager ATen path on the GPU

Workload (BASELINE.json configs[3]): SDXL-shaped UNet (788 wrapped layers,
F1 = 47.81 TFLOP per dense pass at batch 8), bf16 base weights, fp32 adapter parameters under
torch.autocast(bf16) (the kohya regime), LoKr factor 8 full-dim (network_dim 100000) via
lycoris_b200.kohya.create_network, preset "full", per-GPU batch 8, 1024x1024 (latents 128x128),
synthetic N(0,1) latents/context, MSE loss in fp32.  A step is forward + backward (+ the NCCL
adapter-gradient all-reduce when N > 1); there is no optimizer step in the metric.

Prints ONE JSON line (rank 0).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="engine", choices=["engine", "reference", "reference-gpu"])
    ap.add_argument("--batch", type=int, default=8, help="per-GPU batch")
    ap.add_argument("--model", default="sdxl", choices=["sdxl", "sd15", "toy"])
    ap.add_argument("--algo", default="lokr")
    ap.add_argument("--no-graph", action="store_true", help="eager launches instead of a CUDA graph")
    ap.add_argument("--sample-size", type=int, default=0, help="latent side (default: the model's)")
    ap.add_argument("--cpu-seconds", type=float, default=20.0, help="budget of the cpu_baseline leg")
    ap.add_argument("--skip-cpu-baseline", action="store_true")
    ap.add_argument("--kernel-table", default="", help="write a per-kernel CUDA-time table of one eager step here")
    ap.add_argument("--nvtx-step", action="store_true", help="wrap ONE extra eager step in an NVTX range 'lyco_step' (for ncu --nvtx-include)")
    return ap.parse_args()


# ------------------------------------------------------------------------------ clocks
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.rows = []
        self._stop = threading.Event()
        self._t = None

    def _run(self):
        while not self._stop.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                      "-i", str(self.index)], capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.rows.append([c.strip() for c in out.split(",")])
            except Exception:  # noqa: BLE001
                pass
            self._stop.wait(0.2)

    def __enter__(self):
        self._t = threading.Thread(target=self._run, daemon=True)
        self._t.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        self._t.join(timeout=6)

    def summary(self):
        sm, mx, reasons = [], 0, set()
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                mx = max(mx, float(r[1]))
            except (ValueError, IndexError):
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        med = sm[len(sm) // 2] if sm else None
        return {"sm_mhz": med, "sm_max_mhz": mx or None, "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------ workload
def model_cfg(name):
    from workloads.unet_skeleton import SD15, SDXL, TOY

    return {"sdxl": SDXL, "sd15": SD15, "toy": TOY}[name]


def network_args(algo):
    if algo == "lokr":
        return dict(network_dim=100000, network_alpha=1, kw=dict(algo="lokr", factor=8, preset="full"))
    if algo == "locon":
        return dict(network_dim=16, network_alpha=8, kw=dict(algo="locon", conv_dim=8, conv_alpha=8, preset="full"))
    if algo == "loha":
        return dict(network_dim=32, network_alpha=16, kw=dict(algo="loha", conv_dim=16, conv_alpha=8, preset="full"))
    raise KeyError(algo)


NCU_TRAFFIC_DOMINANT = 183.3e6  # bytes per launch (52.8 MB read + 130.5 MB written), see profiles/

ALGO_NOTE = {"lokr": "factor 8, full-dim", "locon": "dim 16 conv_dim 8 alpha 8", "loha": "dim 32 conv_dim 16"}


def perturb_zero_factors(net, seed=1):
    import torch

    g = torch.Generator(device="cpu").manual_seed(seed)
    with torch.no_grad():
        for p in net.parameters():
            if float(p.abs().sum()) == 0.0:
                p.copy_((torch.randn(p.shape, generator=g) * 0.01).to(p.device, p.dtype))


def build_engine_workload(args, device):
    import torch

    import lycoris_b200.kohya as kohya
    from workloads.unet_skeleton import UNetSkeleton, wrapped_layer_flops

    cfg = model_cfg(args.model)
    torch.manual_seed(0)
    # activations travel channels_last (NHWC) — what the TMA-im2col producer and cuDNN both want; the
    # frozen filters stay in PyTorch's [O, C, kh, kw] layout, which is the layout the reference flattens
    unet = UNetSkeleton(cfg).to(device=device, dtype=torch.bfloat16)
    unet.requires_grad_(False)
    unet.train()
    na = network_args(args.algo)
    torch.manual_seed(1)
    net = kohya.create_network(1.0, na["network_dim"], na["network_alpha"], None, None, unet, **na["kw"])
    net.apply_to(None, unet, False, True)
    net.to(device)
    perturb_zero_factors(net, 1)
    net.requires_grad_(True)
    net.train()
    f1, n_layers = wrapped_layer_flops(unet, args.batch, args.sample_size or None)
    return unet, net, f1, n_layers


def make_step(unet, net, static, dp):
    import torch
    import torch.nn.functional as F

    def step():
        if dp is not None:
            dp.zero_grad()
        else:
            for p in net.parameters():
                p.grad = None
        with torch.autocast("cuda", dtype=torch.bfloat16):
            out = unet(static["sample"], static["timesteps"], static["context"], static.get("added_cond"))
        loss = F.mse_loss(out.float(), static["target"].float())
        loss.backward()
        return loss

    return step


def run_engine(args):
    import torch
    import torch.distributed as dist

    from lycoris_b200.engine import _lib
    from lycoris_b200.engine import kernels as K
    from lycoris_b200.engine.ddp import FlatGradAllReduce

    rank = int(os.environ.get("RANK", 0))
    local_rank = int(os.environ.get("LOCAL_RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    torch.cuda.set_device(local_rank)
    device = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=device)
    assert _lib.load().lyco_device_check(local_rank) == 0, _lib.last_error()
    # Everything runs on one non-default stream: autograd binds each parameter's AccumulateGrad node to
    # the stream of its first use, and a node bound to the legacy default stream cannot be captured.
    main_stream = torch.cuda.Stream(device=device)
    torch.cuda.set_stream(main_stream)

    unet, net, f1, n_layers = build_engine_workload(args, device)
    n_params = sum(p.numel() for p in net.parameters())
    dp = FlatGradAllReduce(list(net.parameters()), overlap=True) if world > 1 else None

    cfg = unet.cfg
    host = unet.synthetic_batch(args.batch, "cpu", torch.bfloat16, seed=2 + rank, sample_size=args.sample_size or None)
    host = {k: (v.to(torch.bfloat16) if v.is_floating_point() else v).pin_memory() for k, v in host.items()}
    static = {}
    for k, v in host.items():
        t = v.to(device, non_blocking=True)
        if k in ("sample", "target"):
            t = t.contiguous(memory_format=torch.channels_last)
        static[k] = t
    h2d_bytes = sum(v.numel() * v.element_size() for v in host.values())
    step = make_step(unet, net, static, dp)

    # warm-up (eager) — also primes cuDNN heuristics, TMA descriptors, cached host scalars
    torch.cuda.synchronize()
    for _ in range(max(args.warmup, 3)):
        loss = step()
        if dp is not None:
            dp.allreduce()
            dp.wait()
    torch.cuda.synchronize()

    graph = None
    static_loss = None
    if not args.no_graph:
        try:
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph):
                static_loss = step()
            graph.replay()
            torch.cuda.synchronize()
        except Exception as e:  # noqa: BLE001
            if rank == 0:
                print(f"[bench] CUDA graph capture failed ({type(e).__name__}: {e}); running eager", file=sys.stderr)
            graph = None
            torch.cuda.synchronize()

    def one_step():
        if graph is not None:
            graph.replay()
            out = static_loss
        else:
            out = step()
        if dp is not None:
            dp.allreduce()
            dp.wait()
        return out

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(n, e2e):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        e0.record()
        last = None
        for _ in range(n):
            if e2e:
                for k, v in host.items():
                    static[k].copy_(v, non_blocking=True)
                last = float(one_step().detach())  # device -> host read of the step's result
            else:
                last = one_step()
        e1.record()
        barrier()
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], device=device)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t)
        return ms / n, last

    for _ in range(args.warmup):
        one_step()
    launches0 = _lib.launch_count()
    with ClockSampler(local_rank) as clocks:
        ms_step, _ = timed(args.steps, e2e=False)
        ms_e2e, loss_val = timed(args.steps, e2e=True)
    launches = _lib.launch_count() - launches0
    if graph is not None:
        # graph replays do not pass through the C-ABI; count the launches of one captured step instead
        launches = None

    # instrumented eager steps: CUDA-event pair around every lyco_gemm launch (same stream)
    sink = []
    K.set_gemm_profiler(sink)
    l0 = _lib.launch_count()
    # keep the GPU busy for ~0.6 s first so the host runs ahead of it: the event pairs then bracket kernel
    # execution on the stream, not the GPU waiting for the (slower) eager Python launch path
    torch.cuda._sleep(int(0.6 * 1.9e9))
    step()
    torch.cuda.synchronize()
    per_step_launches = _lib.launch_count() - l0
    K.set_gemm_profiler(None)
    gemm_ms = sum(e0.elapsed_time(e1) for e0, e1, *_ in sink)
    gemm_flops = sum(f for _, _, f, *_ in sink)
    if args.kernel_table and rank == 0:
        by_shape = {}
        for e0_, e1_, fl, M_, N_, K_ in sink:
            a = by_shape.setdefault((M_, N_, K_), [0, 0.0, 0.0])
            a[0] += 1
            a[1] += e0_.elapsed_time(e1_)
            a[2] += fl
        with open(args.kernel_table + ".gemm_shapes", "w") as fh:
            fh.write("lyco_gemm calls of one eager step by (M, N, K): count, total ms (event-bracketed), TFLOP/s\n")
            for (M_, N_, K_), (cnt, ms_, fl) in sorted(by_shape.items(), key=lambda kv: -kv[1][1]):
                fh.write(f"{ms_:9.3f} ms {cnt:5d}  M={M_:7d} N={N_:6d} K={K_:7d}  {fl / ms_ / 1e9:8.1f} TF\n")
    torch.cuda._sleep(int(0.6 * 1.9e9))
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    step()
    e1.record()
    torch.cuda.synchronize()
    eager_ms = e0.elapsed_time(e1)

    gemm_kernel_ms = None
    if rank == 0:
        from torch.profiler import ProfilerActivity, profile, record_function

        from lycoris_b200.engine import ops as _ops

        def _labelled(fn, label):
            def wrapped(*a, **k):
                with record_function(label):
                    return fn(*a, **k)
            return staticmethod(wrapped)

        saved = {}
        for cls in (_ops._AdapterContraction, _ops._MergedContraction):
            saved[cls] = (cls.forward, cls.backward)
            cls.forward = _labelled(cls.forward, "lyco_node_fwd")
            cls.backward = _labelled(cls.backward, "lyco_node_bwd")
        try:
            with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
                step()
                torch.cuda.synchronize()
        finally:
            for cls, (f, b) in saved.items():
                cls.forward, cls.backward = staticmethod(f), staticmethod(b)
        # attribute every kernel to "inside an engine autograd node" or "model side" via the CPU op that launched it
        inside, outside = {}, {}
        for ev in prof.events():
            ks = getattr(ev, "kernels", None)
            if not ks:
                continue
            anc, tag = ev, None
            while anc is not None:
                if anc.name in ("lyco_node_fwd", "lyco_node_bwd"):
                    tag = anc.name
                    break
                anc = anc.cpu_parent
            for k in ks:
                name = k.name.split("<")[0][:70]
                dst = inside if tag else outside
                a = dst.setdefault(name, [0, 0.0])
                a[0] += 1
                a[1] += k.duration
        if args.kernel_table:
            with open(args.kernel_table + ".attribution", "w") as fh:
                for title, d in (("launched inside the engine's autograd nodes", inside),
                                 ("model side (outside)", outside)):
                    tot_d = sum(v[1] for v in d.values())
                    fh.write(f"{title}: {tot_d / 1e3:.2f} ms\n")
                    for name, (cnt, t) in sorted(d.items(), key=lambda kv: -kv[1][1])[:14]:
                        fh.write(f"{t / 1e3:10.3f} ms {cnt:6d}  {name}\n")
        agg = {}
        for ev in prof.events():
            if (ev.device_type is not None and str(ev.device_type).endswith("CUDA") and ev.device_time_total > 0
                    and not ev.name.startswith("lyco_node_")):
                name = ev.name.split("<")[0][:90]
                a = agg.setdefault(name, [0, 0.0])
                a[0] += 1
                a[1] += ev.device_time_total
        tot = sum(v[1] for v in agg.values())
        gemm_kernel_ms = sum(v[1] for k, v in agg.items() if "gemm_sm100_kernel" in k) / 1e3
        if args.kernel_table:
            with open(args.kernel_table, "w") as fh:
                fh.write(f"one eager step, CUDA kernels by total device time (us); sum = {tot / 1e3:.2f} ms\n")
                for name, (cnt, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
                    fh.write(f"{t / 1e3:10.3f} ms {100 * t / tot:5.1f}% {cnt:6d}  {name}\n")
    if args.nvtx_step:
        # start/end (not push/pop) ranges are process-wide: backward kernels are launched from autograd's
        # worker thread and would fall outside a thread-local push/pop range
        torch.cuda.synchronize()
        rid = torch.cuda.nvtx.range_start("lyco_step")
        step()
        torch.cuda.synchronize()
        torch.cuda.nvtx.range_end(rid)

    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:  # noqa: BLE001
        pass
    peak_tf = peaks.get("bf16_tflops_sustained", 1400.0)
    achieved_tf = gemm_flops / (gemm_ms * 1e-3) / 1e12 if gemm_ms > 0 else 0.0

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    steps_per_s = world * 1000.0 / ms_step
    result = {
        "metric": f"SDXL-UNet+{ {'lokr': 'LoKr', 'locon': 'LoCon', 'loha': 'LoHa'}[args.algo] } fwd+bwd steps/sec",
        "value": steps_per_s,
        "unit": "steps/s",
        "n_gpus": world,
        "steps": args.steps,
        "warmup": args.warmup,
        "ms_per_step": ms_step,
        "higher_is_better": True,
        "scaling": "weak",
        "vs_baseline": None,
        "dtype": "bf16",
        "data": "synthetic",
        "impl": "engine",
        "config": {
            "workload": f"{cfg.name}-unet-skeleton + {args.algo} ({ALGO_NOTE[args.algo]}) via lycoris_b200.kohya, preset full, "
                        f"per-GPU batch {args.batch}, latents {args.sample_size or cfg.sample_size}^2, fwd+bwd, "
                        "bf16 base / fp32 adapter / autocast",
            "wrapped_layers": n_layers,
            "adapter_params": n_params,
            "global_batch": args.batch * world,
            "parallelism": f"dp{world}",
            "cuda_graph": graph is not None,
            "l2": "inputs+weights+activations per step (>10 GB) exceed the 126 MB L2",
            "F1_tflop_per_dense_pass": f1 / 1e12,
        },
        "e2e": {
            "value": world * 1000.0 / ms_e2e, "unit": "steps/s", "ms_per_step": ms_e2e,
            "h2d_bytes_per_step": h2d_bytes, "d2h_bytes_per_step": 4, "loss": loss_val,
        },
        "gpu_launches": per_step_launches * args.steps * 2,
        "lyco_launches_per_step": per_step_launches,
        "roofline": {
            "bound": "tensor", "kernel": "gemm_sm100_kernel (fwd / dgrad / wgrad of the wrapped Linear layers)",
            "achieved": achieved_tf, "peak": peak_tf, "unit": "TFLOP/s", "frac": achieved_tf / peak_tf,
            "peak_source": "MEASURED_PEAKS.json bf16_tflops_sustained (of measured)" if peaks else "fallback 1400",
            # dram__bytes_read+write per launch of the step's dominant GEMM (GEGLU proj forward, M=8192 N=10240
            # K=1280) from the committed `ncu --set full` capture profiles/r01_ncu_full_summary.txt
            "traffic": NCU_TRAFFIC_DOMINANT if args.model == "sdxl" else None,
            "traffic_shape": "M=8192 N=10240 K=1280 fwd: algorithmic 2*(MK+NK+MN) = 215.0e6 B" if args.model == "sdxl" else None,
            "gemm_launches_per_step": len(sink), "gemm_ms_per_step": gemm_ms, "gemm_share_of_eager_step": gemm_ms / eager_ms, "eager_step_ms_gpu_bound": eager_ms,
            "algorithmic_tflop_per_step": gemm_flops / 1e12,
            # the event brackets above also contain the wgrad memsets and ~5 us of stream front-end gap per launch;
            # the same launches by CUPTI kernel duration (torch.profiler over one eager step):
            "gemm_kernel_ms_per_step_cupti": gemm_kernel_ms,
            "achieved_cupti": (gemm_flops / (gemm_kernel_ms * 1e-3) / 1e12) if gemm_kernel_ms else None,
        },
        "clocks": clocks.summary(),
    }
    if not args.skip_cpu_baseline and world == 1:
        result["cpu_baseline"] = cpu_reference(args, budget_s=args.cpu_seconds)
    print(json.dumps(result), flush=True)
    if world > 1:
        dist.destroy_process_group()


# ------------------------------------------------------------------ reference arms
def distinct_wrapped_shapes(args):
    """[(kind, N, K, ksize, stride, M_per_sample_or_tokens, count)] of the wrapped layers."""
    import torch
    import torch.nn as nn

    from workloads.unet_skeleton import UNetSkeleton

    cfg = model_cfg(args.model)
    with torch.device("meta"):
        unet = UNetSkeleton(cfg)
    targets = ("Transformer2DModel", "ResnetBlock2D", "Downsample2D", "Upsample2D")
    layers = {}
    for _, mod in unet.named_modules():
        if mod.__class__.__name__ in targets:
            for _, sub in mod.named_modules():
                if isinstance(sub, (nn.Linear, nn.Conv2d)):
                    layers[id(sub)] = sub
    shapes = {}

    def hook(m, inp, out):
        x = inp[0]
        if isinstance(m, nn.Linear):
            key = ("linear", m.out_features, m.in_features, 1, 1, x.numel() // x.shape[-1] // args.batch, 0)
        else:
            key = ("conv", m.out_channels, m.in_channels, m.kernel_size[0], m.stride[0], x.shape[-1], m.padding[0])
        shapes[key] = shapes.get(key, 0) + 1

    hs = [m.register_forward_hook(hook) for m in layers.values()]
    s = args.sample_size or cfg.sample_size
    with torch.no_grad():
        b = {
            "sample": torch.zeros(args.batch, cfg.in_channels, s, s, device="meta"),
            "t": torch.zeros(args.batch, device="meta", dtype=torch.long),
            "ctx": torch.zeros(args.batch, cfg.context_len, cfg.cross_attention_dim, device="meta"),
            "add": torch.zeros(args.batch, cfg.addition_embed_dim, device="meta") if cfg.addition_embed_dim else None,
        }
        unet(b["sample"], b["t"], b["ctx"], b["add"])
    for h in hs:
        h.remove()
    return shapes


def cpu_reference(args, budget_s=20.0):
    """The reference's CPU path (oracle port of lycoris/modules/lokr.py forward, autograd backward),
    timed on this box's host cores on a BOUNDED sample: every distinct wrapped-layer shape of the
    workload is run fwd+bwd at a reduced number of rows / reduced spatial size, scaled linearly to
    the full M and multiplied by its multiplicity; un-wrapped ops (attention, norms) are not counted,
    which flatters the CPU.  fp32 (the reference's CPU-runnable regime)."""
    import torch

    from oracle import lyco_oracle as O

    threads = torch.get_num_threads()
    shapes = distinct_wrapped_shapes(args)
    total_full = 0.0
    t_start = time.time()
    per_shape_budget = budget_s / max(1, len(shapes))
    sampled = 0
    for (kind, N, Kd, ks, stride, m, pad), count in shapes.items():
        torch.manual_seed(0)
        (a, b), (c, d) = O.factorization(N, 8), O.factorization(Kd, 8)
        if kind == "linear":
            rows_full = m * args.batch
            rows = max(8, min(rows_full, 256))
            x = torch.randn(rows, Kd)
            W = torch.randn(N, Kd) * 0.02
            p = {"lokr_w1": torch.randn(a, c) * 0.1, "lokr_w2": torch.randn(b, d) * 0.02}
            conv = None
            scale_up = rows_full / rows
        else:
            side_full = m
            side = max(8, min(side_full, 16))
            x = torch.randn(1, Kd, side, side)
            W = torch.randn(N, Kd, ks, ks) * 0.02
            p = {"lokr_w1": torch.randn(a, c) * 0.1, "lokr_w2": torch.randn(b, d, ks, ks) * 0.02}
            conv = dict(stride=(stride, stride), padding=(pad, pad), dilation=(1, 1), groups=1)
            scale_up = args.batch * (side_full / side) ** 2
        bias = torch.zeros(N)
        cfg = {"scale": 1.0, "multiplier": 1.0}
        # time: at least one rep, stop at this shape's share of the budget
        reps, t_acc = 0, 0.0
        while reps < 1 or (t_acc < per_shape_budget * 0.5 and reps < 3):
            t0 = time.perf_counter()
            xx = x.clone().requires_grad_(True)
            leaves = {k: v.clone().requires_grad_(True) for k, v in p.items()}
            out = O.layer_forward("lokr", xx, W, bias, leaves, cfg, conv)
            out.float().pow(2).mean().backward()
            t_acc += time.perf_counter() - t0
            reps += 1
        total_full += (t_acc / reps) * scale_up * count
        sampled += 1
    steps_per_s = 1.0 / total_full if total_full > 0 else 0.0
    return {
        "value": steps_per_s, "unit": "steps/s", "cores": threads, "kind": "port",
        "sample": f"oracle (torch-CPU restatement of the reference path), fp32, {sampled} distinct wrapped-layer shapes "
                  f"timed fwd+bwd at <=256 rows / <=16x16 spatial, scaled linearly to batch {args.batch} and multiplied "
                  f"by multiplicity; un-wrapped ops excluded; extrapolated; {time.time() - t_start:.1f}s of CPU work",
        "extrapolated_s_per_step": total_full,
        "host_cpus": os.cpu_count(),
    }


def run_reference_cpu(args):
    rank = int(os.environ.get("RANK", 0))
    if rank != 0:
        return
    cfg = model_cfg(args.model)
    vals = []
    for _ in range(max(1, min(args.warmup, 1))):
        cpu_reference(args, budget_s=min(5.0, args.cpu_seconds))
    for _ in range(max(1, args.steps if args.steps < 3 else 3)):
        vals.append(cpu_reference(args, budget_s=args.cpu_seconds))
    best = max(vals, key=lambda r: r["value"])
    result = {
        "metric": "SDXL-UNet+LoKr fwd+bwd steps/sec", "value": best["value"], "unit": "steps/s",
        "n_gpus": int(os.environ.get("WORLD_SIZE", 1)), "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1000.0 * best["extrapolated_s_per_step"], "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic", "impl": "reference",
        # the engine arm's workload, timed on the host cores through the oracle port of the reference path
        "config": {"workload": f"{cfg.name}-unet-skeleton + {args.algo} ({ALGO_NOTE[args.algo]}), preset full, "
                               f"per-GPU batch {args.batch}, latents {args.sample_size or cfg.sample_size}^2, fwd+bwd, "
                               "reference path (oracle port) on the host CPU, fp32",
                   "global_batch": args.batch, "parallelism": "cpu", "sample": best["sample"]},
        "cpu_baseline": best,
        "e2e": {"value": best["value"], "unit": "steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(result), flush=True)


def run_reference_gpu(args):
    """Informational: what the reference's eager path costs on this GPU.  Uses the unmodified
    reference if it was pip-installed into baseline/_ref, else the oracle's per-layer forward
    (same ATen calls) patched onto every wrapped layer."""
    import torch
    import torch.nn.functional as F

    torch.cuda.set_device(0)
    device = torch.device("cuda", 0)
    from workloads.unet_skeleton import UNetSkeleton

    cfg = model_cfg(args.model)
    torch.manual_seed(0)
    unet = UNetSkeleton(cfg).to(device=device, dtype=torch.bfloat16).to(memory_format=torch.channels_last)
    unet.requires_grad_(False)
    unet.train()
    na = network_args(args.algo)
    ref_dir = os.path.join(ROOT, "baseline", "_ref")
    source = "oracle-patched layers"
    if os.path.isdir(os.path.join(ref_dir, "lycoris")):
        sys.path.insert(0, ref_dir)
        import lycoris.kohya as ref_kohya

        torch.manual_seed(1)
        net = ref_kohya.create_network(1.0, na["network_dim"], na["network_alpha"], None, None, unet, **na["kw"])
        net.apply_to(None, unet, False, True)
        net.to(device)
        source = "unmodified reference from baseline/_ref"
    else:
        import lycoris_b200.kohya as kohya
        from oracle import lyco_oracle as O

        torch.manual_seed(1)
        net = kohya.create_network(1.0, na["network_dim"], na["network_alpha"], None, None, unet, **na["kw"])
        for lora in net.loras:
            net.add_module(lora.lora_name, lora)
        net.to(device)
        for lora in net.loras:  # patch the oracle's forward instead of the engine's
            org = lora.org_module[0]
            conv = None
            if lora.module_type.startswith("conv"):
                conv = dict(stride=org.stride, padding=org.padding, dilation=org.dilation, groups=org.groups)

            def fwd(x, _l=lora, _o=org, _c=conv):
                p = {k: v for k, v in _l.named_parameters()}
                return O.layer_forward("lokr", x, _o.weight, _o.bias, p, {"scale": _l.scale, "multiplier": 1.0}, _c)

            org.forward = fwd
    perturb_zero_factors(net, 1)
    net.requires_grad_(True)
    batch = unet.synthetic_batch(args.batch, "cpu", torch.bfloat16, seed=2, sample_size=args.sample_size or None)
    st = {k: (v.to(torch.bfloat16) if v.is_floating_point() else v).to(device) for k, v in batch.items()}
    for k in ("sample", "target"):
        st[k] = st[k].contiguous(memory_format=torch.channels_last)

    def step():
        for p in net.parameters():
            p.grad = None
        with torch.autocast("cuda", dtype=torch.bfloat16):
            out = unet(st["sample"], st["timesteps"], st["context"], st.get("added_cond"))
        F.mse_loss(out.float(), st["target"].float()).backward()

    for _ in range(max(3, args.warmup)):
        step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        step()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / args.steps
    print(json.dumps({"metric": "SDXL-UNet+LoKr fwd+bwd steps/sec", "impl": "reference-gpu", "source": source,
                      "value": 1000.0 / ms, "unit": "steps/s", "ms_per_step": ms, "n_gpus": 1, "steps": args.steps,
                      "warmup": args.warmup, "dtype": "bf16", "data": "synthetic",
                      "config": {"workload": f"{cfg.name} + {args.algo}, batch {args.batch}, eager ATen path"}}), flush=True)


def main():
    args = parse()
    if args.impl == "reference":
        run_reference_cpu(args)
    elif args.impl == "reference-gpu":
        run_reference_gpu(args)
    else:
        run_engine(args)


if __name__ == "__main__":
    main()
