#!/usr/bin/env python
"""bench.py — LyCORIS adapter fwd+bwd step rate on N x B200 (default: SDXL-UNet + LoKr factor 8 full-dim).

    python bench.py --gpus N --steps K --warmup W              # this engine (default arm), BASELINE cfg #4
    python bench.py --config cfg2|cfg3|cfg4|cfg5 ...           # the other BASELINE.json configs
    python bench.py --impl reference ...                       # the reference's own CPU path on the host cores
    python bench.py --impl reference-gpu ...                   # the UNMODIFIED reference, PyTorch-eager, on one GPU

Workloads (BASELINE.json configs): SDXL- / SD1.5-shaped UNet skeleton, bf16 base weights, fp32
adapter parameters under torch.autocast(bf16) (the kohya regime), adapters created through
``<package>.kohya.create_network`` exactly as kohya sd-scripts does, synthetic N(0,1) latents / context,
MSE loss in fp32.  A step is forward + backward (+ the NCCL adapter-gradient all-reduce, overlapped with
backward, when N > 1); there is no optimizer step in the metric.

The engine arm at N = 1 also times, in the same process and on the same model / parameters / inputs, the
unmodified reference from baseline/_ref through ITS public API (``lycoris.kohya``), PyTorch-eager — the
denominator BASELINE.json's >=4x target names — cross-checks the two losses, and times the reference's CPU
path on the host cores on a bounded sample (``cpu_baseline``).

Prints ONE JSON line (rank 0).
"""
import argparse
import json
import logging
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
REF_DIR = os.path.join(ROOT, "baseline", "_ref")

CFG5_PRESET = os.path.join(ROOT, "workloads", "cfg5_mixed_preset.toml")

# BASELINE.json configs[1..4]; per-GPU batch except cfg5 whose batch 16 is GLOBAL (split over the ranks)
CONFIGS = {
    "cfg2": dict(model="sd15", batch=4, label="LoCon", note="dim 16 conv_dim 8 alpha 8",
                 dim=16, alpha=8, kw=dict(algo="locon", conv_dim=8, conv_alpha=8, preset="full")),
    "cfg3": dict(model="sdxl", batch=8, label="LoHa", note="dim 32 conv_dim 16",
                 dim=32, alpha=16, kw=dict(algo="loha", conv_dim=16, conv_alpha=8, preset="full")),
    "cfg4": dict(model="sdxl", batch=8, label="LoKr", note="factor 8, full-dim",
                 dim=100000, alpha=1, kw=dict(algo="lokr", factor=8, preset="full")),
    "cfg5": dict(model="sdxl", batch=16, global_batch=True, label="mixed(LoCon+LoHa+LoKr+IA3)",
                 note="locon d16/c8 on ResnetBlock2D + samplers, lokr f8 full on FeedForward, loha d16 on attention "
                      "q/out + proj, ia3 on to_k/to_v (workloads/cfg5_mixed_preset.toml)",
                 dim=16, alpha=8, kw=dict(algo="locon", conv_dim=8, conv_alpha=4, preset=CFG5_PRESET)),
}
ALGO_TO_CFG = {"lokr": "cfg4", "loha": "cfg3", "locon": "cfg2"}


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="engine", choices=["engine", "reference", "reference-gpu"])
    ap.add_argument("--config", default="", choices=["", "cfg2", "cfg3", "cfg4", "cfg5"],
                    help="BASELINE.json config (default cfg4 = SDXL + LoKr f8 full-dim, per-GPU batch 8)")
    ap.add_argument("--batch", type=int, default=0, help="per-GPU batch override")
    ap.add_argument("--model", default="", choices=["", "sdxl", "sd15", "toy"], help="model override")
    ap.add_argument("--algo", default="", help="shorthand: lokr|loha|locon pick the adapter of cfg4|cfg3|cfg2")
    ap.add_argument("--no-graph", action="store_true", help="eager launches instead of a CUDA graph")
    ap.add_argument("--sample-size", type=int, default=0, help="latent side (default: the model's)")
    ap.add_argument("--cpu-seconds", type=float, default=20.0, help="budget of the cpu_baseline leg")
    ap.add_argument("--skip-cpu-baseline", action="store_true")
    ap.add_argument("--skip-gpu-reference", action="store_true", help="do not time the unmodified reference on the GPU")
    ap.add_argument("--ref-steps", type=int, default=20, help="timed steps of the GPU-eager reference leg")
    ap.add_argument("--kernel-table", default="", help="write a per-kernel CUDA-time table of one eager step here")
    ap.add_argument("--nvtx-step", action="store_true", help="wrap ONE extra eager step in an NVTX range 'lyco_step' (for ncu --nvtx-include)")
    args = ap.parse_args()
    name = args.config or ALGO_TO_CFG.get(args.algo, "cfg4")
    wl = dict(CONFIGS[name])
    wl["name"] = name
    if args.algo and not args.config:
        # --algo on the default model: the adapter settings of that config on the SDXL model (round-1 behaviour)
        wl["model"], wl["batch"] = "sdxl", 8
    if args.model:
        wl["model"] = args.model
    world = int(os.environ.get("WORLD_SIZE", 1))
    if args.batch:
        wl["batch"] = args.batch
        wl["global_batch"] = False
    elif wl.get("global_batch"):
        # cfg5's batch 16 is the GLOBAL batch of an 8-GPU run (2 per GPU).  One 180 GB GPU cannot hold 16 samples of this
        # workload (≈8.8 GB of saved activations per sample at 1024², measured: OOM at 178 GB during the forward, in the
        # engine arm and — with its extra delta weights — in the reference arm alike), so N = 1 runs 8.
        wl["batch"] = max(1, min(wl["batch"] // world, 8))
    args.wl = wl
    return args


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


def perturb_zero_factors(net, seed=1):
    import torch

    g = torch.Generator(device="cpu").manual_seed(seed)
    with torch.no_grad():
        for p in net.parameters():
            if float(p.abs().sum()) == 0.0:
                p.copy_((torch.randn(p.shape, generator=g) * 0.01).to(p.device, p.dtype))


def create_network(kohya_mod, wl, unet):
    """The call kohya sd-scripts makes: ``network_module.create_network(multiplier, dim, alpha, vae, te, unet, **kw)``."""
    import torch

    torch.manual_seed(1)
    return kohya_mod.create_network(1.0, wl["dim"], wl["alpha"], None, None, unet, **wl["kw"])


def import_reference():
    """The UNMODIFIED reference package vendored by ``__graft_entry__.build()`` (pip install --target
    baseline/_ref).  Returns (lycoris.kohya module, notes) or (None, why)."""
    if not os.path.isdir(os.path.join(REF_DIR, "lycoris")):
        return None, "baseline/_ref is empty (run __graft_entry__.build() where /root/reference exists)"
    if REF_DIR not in sys.path:
        sys.path.insert(0, REF_DIR)
    try:
        import lycoris.kohya as ref_kohya
        import lycoris.wrapper as ref_wrapper
        from lycoris.modules.ia3 import IA3Module
    except Exception as e:  # noqa: BLE001
        return None, f"import failed: {type(e).__name__}: {e}"
    notes = []
    logging.getLogger("LyCORIS").setLevel(logging.ERROR)  # the reference's logging module resets it to INFO at import
    if "ia3" not in ref_wrapper.network_module_dict:
        # upstream omission (lycoris/wrapper.py:45-55): the class exists but is not registered, so any preset that
        # names algo="ia3" raises KeyError.  Registering the reference's OWN class is the one change made.
        ref_wrapper.network_module_dict["ia3"] = IA3Module
        notes.append("registered the reference's own IA3Module under 'ia3' in network_module_dict (upstream omission)")
    return ref_kohya, notes


def build_engine_workload(args, device):
    import torch

    import lycoris_b200.kohya as kohya
    from workloads.unet_skeleton import UNetSkeleton, wrapped_layer_flops

    wl = args.wl
    cfg = model_cfg(wl["model"])
    torch.manual_seed(0)
    # activations travel channels_last (NHWC) — what the TMA-im2col producer and cuDNN both want; the
    # frozen filters stay in PyTorch's [O, C, kh, kw] layout, which is the layout the reference flattens
    unet = UNetSkeleton(cfg).to(device=device, dtype=torch.bfloat16)
    unet.requires_grad_(False)
    unet.train()
    net = create_network(kohya, wl, unet)
    net.apply_to(None, unet, False, True)
    net.to(device)
    perturb_zero_factors(net, 1)
    net.requires_grad_(True)
    net.train()
    f1, n_layers = wrapped_layer_flops(unet, wl["batch"], args.sample_size or None)
    return unet, net, f1, n_layers


def algorithmic_flops(net, rows_of):
    """F_alg = sum over wrapped layers of c * 2*M*N*K' + F_side, c = 2 (LoCon, LoKr, IA3, DyLoRA)
    or 3 (LoHa); F_side: LoCon 3*2*M*r*(N+K'), LoKr 3*2*M*(uq*vq*vp + vp*uq*up), LoHa 8*2*N*K'*r."""
    total = 0.0
    for lora in net.loras:
        M = rows_of.get(lora.lora_name)
        if M is None:
            continue
        N = lora.shape[0]
        Kp = 1
        for s in lora.shape[1:]:
            Kp *= s
        kind = type(lora).__name__
        dense = 2.0 * M * N * Kp
        if kind == "LohaModule":
            total += 3 * dense + 8 * 2.0 * N * Kp * lora.lora_dim
        elif kind == "LokrModule":
            w1 = lora.lokr_w1 if lora.use_w1 else None
            up, uq = (w1.shape if w1 is not None else (lora.lokr_w1_a.shape[0], lora.lokr_w1_b.shape[1]))
            vp, vq = N // up, Kp // uq
            total += 2 * dense + 3 * 2.0 * M * (uq * vq * vp + vp * uq * up)
        elif kind in ("LoConModule", "DyLoraModule"):
            total += 2 * dense + 3 * 2.0 * M * lora.lora_dim * (N + Kp)
        else:
            total += 2 * dense
    return total


def layer_rows(unet, net, batch, sample_size=None):
    """{lora_name: M} — rows (batch x positions / tokens) each wrapped layer contracts over, from a no-grad
    forward with hooks on the base layers (run BEFORE timing; the tensors are tiny bookkeeping)."""
    import torch

    rows = {}
    hooks = []
    for lora in net.loras:
        org = lora.org_module[0]

        def hook(m, inp, out, _n=lora.lora_name):
            o = out
            rows[_n] = o.numel() // o.shape[1] if o.dim() == 4 else o.numel() // o.shape[-1]

        hooks.append(org.register_forward_hook(hook))
    return rows, hooks


def make_step(unet, net, static, dp, comm=True):
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
        if dp is not None and comm:
            dp.allreduce()  # buckets not yet issued from inside backward
            dp.wait()       # join the comm stream (inside the captured graph)
        return loss

    return step


def ncu_traffic(wl_name):
    """dram bytes per launch of the dominant kernel from a COMMITTED ncu capture (profiles/*traffic*.json written by
    tools/ncu_summarize.py); None when no capture for this workload is committed."""
    import glob

    best = None
    for path in sorted(glob.glob(os.path.join(ROOT, "profiles", "r*_ncu_traffic_*.json"))):
        try:
            rec = json.load(open(path))
        except Exception:  # noqa: BLE001
            continue
        if rec.get("workload", "cfg4") == wl_name:
            best = dict(rec, file=os.path.relpath(path, ROOT))
    return best


def run_engine(args):
    import torch
    import torch.distributed as dist

    from lycoris_b200.engine import _lib
    from lycoris_b200.engine import kernels as K
    from lycoris_b200.engine.ddp import FlatGradAllReduce

    wl = args.wl
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
    algo_table = {}
    for lora in net.loras:
        algo_table[type(lora).__name__] = algo_table.get(type(lora).__name__, 0) + 1
    dp = FlatGradAllReduce(list(net.parameters()), overlap="backward") if world > 1 else None

    cfg = unet.cfg
    host = unet.synthetic_batch(wl["batch"], "cpu", torch.bfloat16, seed=2 + rank, sample_size=args.sample_size or None)
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
    rows_of, hooks = layer_rows(unet, net, wl["batch"])
    loss = step()
    for h in hooks:
        h.remove()
    for _ in range(max(args.warmup, 3) - 1):
        loss = step()
    torch.cuda.synchronize()
    f_alg = algorithmic_flops(net, rows_of)

    def capture(fn):
        try:
            torch.cuda.synchronize()
            torch.cuda.empty_cache()  # the eager warm-up's cached blocks: the graph gets its own pool (cfg5: batch 16)
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                out = fn()
            g.replay()
            torch.cuda.synchronize()
            return g, out
        except Exception as e:  # noqa: BLE001
            if rank == 0:
                print(f"[bench] CUDA graph capture failed ({type(e).__name__}: {e}); running eager", file=sys.stderr)
            torch.cuda.synchronize()
            return None, None

    graph, static_loss = (None, None) if args.no_graph else capture(step)
    comm_mode = "none" if dp is None else ("in-graph" if graph is not None else "eager")
    if dp is not None and graph is None and not args.no_graph:
        # the step with its NCCL kernels could not be captured: capture the compute alone (hooks disarmed) and issue
        # the bucketed all-reduce after each replay (exposed, as in round 1, but still a graph-launched step)
        dp._armed = False
        step_nc0 = make_step(unet, net, static, dp, comm=False)
        graph, static_loss = capture(step_nc0)
        if graph is not None:
            comm_mode = "after-graph"
        else:
            dp._armed = True

    def one_step():
        if graph is not None:
            graph.replay()
            if comm_mode == "after-graph":
                dp.allreduce()
                dp.wait()
            return static_loss
        return step()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(n, e2e, fn=None):
        fn = fn or one_step
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        e0.record()
        last = None
        for _ in range(n):
            if e2e:
                for k, v in host.items():
                    static[k].copy_(v, non_blocking=True)
                last = float(fn().detach())  # device -> host read of the step's result
            else:
                last = fn()
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
    with ClockSampler(local_rank) as clocks:
        ms_step, _ = timed(args.steps, e2e=False)
        ms_e2e, loss_val = timed(args.steps, e2e=True)

    # exposed communication: the same step captured WITHOUT the collective, timed the same way (fewer steps)
    allreduce_exposed_ms = None
    buckets_overlapped = None
    if dp is not None:
        buckets_overlapped = [dp.buckets_overlapped, len(dp.buckets)]
        dp._armed = False
        step_nc = make_step(unet, net, static, dp, comm=False)
        g_nc, loss_nc = (None, None) if args.no_graph else capture(step_nc)
        fn_nc = (lambda: (g_nc.replay(), loss_nc)[1]) if g_nc is not None else step_nc
        for _ in range(2):
            fn_nc()
        ms_nc, _ = timed(max(3, min(args.steps, 8)), e2e=False, fn=fn_nc)
        allreduce_exposed_ms = ms_step - ms_nc
        del g_nc
        step = step_nc  # the instrumented eager steps below run without the collective (rank-local)

    # instrumented eager steps: CUDA-event pair around every lyco_gemm launch (same stream)
    # Pass 1 (no brackets kept): how long the HOST needs for one eager step.  Pass 2: before every 8th bracket the
    # stream first spins (outside the brackets) for 1.5x the host time of 8 GEMM calls and everything between them, so
    # the host is ahead of the GPU at every bracket and the event pairs hold kernel time, not the GPU waiting for the
    # eager Python launch path (on a slow or busy host the brackets otherwise count host time: 136 ms instead of 95).
    K.set_gemm_profiler([])
    t_host0 = time.perf_counter()
    step()
    host_s = time.perf_counter() - t_host0
    torch.cuda.synchronize()
    n_brackets = max(1, len(K._gemm_profile))
    refill_every = 8
    refill_s = min(0.025, 1.5 * refill_every * host_s / n_brackets)
    sink = []
    K.set_gemm_profiler(sink, refill_every=refill_every, refill_cycles=int(refill_s * 1.9e9))
    l0 = _lib.launch_count()
    torch.cuda._sleep(int(0.6 * 1.9e9))
    step()
    torch.cuda.synchronize()
    per_step_launches = _lib.launch_count() - l0
    K.set_gemm_profiler(None)
    # split the launches by arithmetic intensity: contractions above the ridge (peak FLOP/s / peak B/s) are bound by the
    # tensor pipe — the dominant kernel class, `roofline` — the skinny structured-gradient contractions below it are
    # HBM-bound and are reported against the copy bandwidth instead (`roofline_hbm_gemm`)
    try:
        _pk = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:  # noqa: BLE001
        _pk = {}
    ridge = _pk.get("bf16_tflops_sustained", 1400.0) * 1e12 / (_pk.get("hbm_gbs", 6650.0) * 1e9)
    dense = [r for r in sink if r[2] / r[6] >= ridge]
    skinny = [r for r in sink if r[2] / r[6] < ridge]
    gemm_ms = sum(r[0].elapsed_time(r[1]) for r in dense)
    gemm_flops = sum(r[2] for r in dense)
    skinny_ms = sum(r[0].elapsed_time(r[1]) for r in skinny)
    skinny_bytes = sum(r[6] for r in skinny)
    skinny_flops = sum(r[2] for r in skinny)
    if args.kernel_table and rank == 0:
        by_shape = {}
        for e0_, e1_, fl, M_, N_, K_, _b in sink:
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
    kernel_ms = {}
    if rank == 0:
        gemm_kernel_ms, kernel_ms = profile_one_step(args, step)
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
        graph = static_loss = step = one_step = None  # drop the captured graphs (they pin the NCCL communicator)
        shutdown_process_group(world)
        return
    steps_per_s = world * 1000.0 / ms_step
    traffic = ncu_traffic(wl["name"])
    result = {
        "metric": f"{cfg.name.upper()}-UNet+{wl['label']} fwd+bwd steps/sec",
        "value": steps_per_s,
        "unit": "steps/s",
        "n_gpus": world,
        "steps": args.steps,
        "warmup": args.warmup,
        "ms_per_step": ms_step,
        "higher_is_better": True,
        "scaling": "strong" if wl.get("global_batch") else "weak",
        "vs_baseline": None,
        "dtype": "bf16",
        "data": "synthetic",
        "impl": "engine",
        "config": {
            "workload": f"{wl['name']}: {cfg.name}-unet-skeleton + {wl['label']} ({wl['note']}) via lycoris_b200.kohya."
                        f"create_network, per-GPU batch {wl['batch']}, latents {args.sample_size or cfg.sample_size}^2, "
                        "fwd+bwd, bf16 base / fp32 adapter / autocast",
            "wrapped_layers": n_layers,
            "adapters": algo_table,
            "adapter_params": n_params,
            "global_batch": wl["batch"] * world,
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
            "bound": "tensor", "kernel": "gemm_sm100_kernel (fwd / dgrad / wgrad contractions of the wrapped Linear layers)",
            "achieved": achieved_tf, "peak": peak_tf, "unit": "TFLOP/s", "frac": achieved_tf / peak_tf,
            "peak_source": "MEASURED_PEAKS.json bf16_tflops_sustained (of measured)" if peaks else "fallback 1400 (of fallback)",
            # dram__bytes_read+write per launch of the dominant GEMM from a committed `ncu --set full` capture
            "traffic": traffic["dram_bytes_per_launch"] if traffic else None,
            "traffic_source": traffic,
            "gemm_launches_per_step": len(dense), "gemm_ms_per_step": gemm_ms, "gemm_share_of_eager_step": gemm_ms / eager_ms,
            "ridge_flop_per_byte": ridge,
            "eager_step_ms_gpu_bound": eager_ms,
            # FLOPs the GEMM launches EXECUTE (every contraction the engine runs, structured or dense)
            "executed_tflop_per_step": gemm_flops / 1e12,
            # the event brackets above also contain the split-K memsets and the stream front-end gap per launch;
            # the same launches by CUPTI kernel duration (torch.profiler over one eager step):
            "gemm_kernel_ms_per_step_cupti": gemm_kernel_ms,
            # all event brackets (tensor-bound + HBM-bound launches) over the CUPTI kernel time of the same launches:
            # ~1.1 when the brackets hold kernels only, well above when the eager step was host-bound
            "event_ms_over_cupti_ms": ((gemm_ms + skinny_ms) / gemm_kernel_ms) if gemm_kernel_ms else None,
            # host time of one eager step and the spin placed in front of every 8th bracket so that the host stays ahead
            "host_eager_step_ms": host_s * 1e3, "bracket_refill_ms": refill_s * 1e3,
            # (CUPTI totals cannot be split by shape: ALL gemm_sm100_kernel launches, tensor-bound and HBM-bound alike)
            "achieved_cupti_all_gemm_launches": ((gemm_flops + skinny_flops) / (gemm_kernel_ms * 1e-3) / 1e12) if gemm_kernel_ms else None,
            # ALGORITHMIC work of the adapter path (c*F1 + F_side, c = 2 or 3 — never the
            # reference's redundant 5*F1) over the WHOLE step time, model-side ops included
            "algorithmic": {
                "tflop_per_step": f_alg / 1e12, "tflops": f_alg / (ms_step * 1e-3) / 1e12,
                "frac_of_peak": f_alg / (ms_step * 1e-3) / 1e12 / peak_tf,
                "note": "F_alg / whole-step time; the step also holds attention, norms and elementwise ops of the model",
            },
            "engine_kernel_ms_per_step_cupti": kernel_ms,
        },
        "clocks": clocks.summary(),
    }
    if skinny:
        hbm = peaks.get("hbm_gbs", 6650.0)
        result["roofline_hbm_gemm"] = {
            "bound": "hbm", "kernel": "gemm_sm100_kernel launches below the ridge (structured LoKr g_w2 / Q contractions, rank-r "
                                      "products, M = 8 time-embedding layers)",
            "achieved": skinny_bytes / (skinny_ms * 1e-3) / 1e9, "peak": hbm, "unit": "GB/s",
            "frac": skinny_bytes / (skinny_ms * 1e-3) / 1e9 / hbm,
            "peak_source": "MEASURED_PEAKS.json hbm_gbs (of measured)" if peaks else "fallback 6650 (of fallback)",
            "launches_per_step": len(skinny), "ms_per_step": skinny_ms, "algorithmic_bytes_per_step": skinny_bytes,
            "executed_tflop_per_step": skinny_flops / 1e12,
        }
    if dp is not None:
        result["allreduce_exposed_ms"] = allreduce_exposed_ms
        result["allreduce"] = {"buckets_issued_inside_backward": buckets_overlapped[0], "buckets": buckets_overlapped[1],
                               "elements": dp.num_elements, "mode": dp.mode, "collectives": comm_mode}
    if world == 1 and not args.skip_gpu_reference:
        del graph
        graph = None
        static_loss = None
        torch.cuda.synchronize()
        torch.cuda.empty_cache()
        result["gpu_eager_reference"] = gpu_eager_reference(args, unet, net, static, loss_val)
        ref = result["gpu_eager_reference"]
        if ref.get("ms_per_step"):
            result["speedup_vs_gpu_eager"] = ref["ms_per_step"] / ms_step
    if not args.skip_cpu_baseline and world == 1:
        result["cpu_baseline"] = cpu_reference(args, budget_s=args.cpu_seconds, reps=3)
    print(json.dumps(result), flush=True)
    graph = static_loss = step = one_step = None
    shutdown_process_group(world)
    ref = result.get("gpu_eager_reference") or {}
    if ref.get("loss_check") == "FAILED":
        # the line above is still printed (with loss_check = FAILED); a parity failure at the headline config must not
        # look like a successful run
        print(f"[bench] PARITY FAILURE: engine loss {ref['engine_loss']} vs reference loss {ref['loss']} "
              f"(relative difference {ref['loss_rel_diff']:.3e} > 2e-2)", file=sys.stderr)
        sys.exit(3)


def shutdown_process_group(world):
    """Leave the NCCL group without hanging: ncclCommDestroy waits for every CUDA graph that captured the communicator
    (the step graph holds the bucketed all-reduces), so the graphs must be gone first; a watchdog ends the process if
    the teardown still blocks (the JSON line is already printed and flushed by then)."""
    if world <= 1:
        return
    import gc

    import torch
    import torch.distributed as dist

    gc.collect()
    torch.cuda.synchronize()
    sys.stdout.flush()
    sys.stderr.flush()
    dog = threading.Timer(20.0, lambda: os._exit(0))
    dog.daemon = True
    dog.start()
    try:
        dist.destroy_process_group()
    except Exception:  # noqa: BLE001
        pass
    dog.cancel()


def profile_one_step(args, step):
    """torch.profiler (CUPTI) over one eager step: GEMM kernel time, per-engine-kernel totals, optional tables."""
    import torch
    from torch.profiler import ProfilerActivity, profile, record_function

    from lycoris_b200.engine import ops as _ops

    def _labelled(fn, label):
        def wrapped(*a, **k):
            with record_function(label):
                return fn(*a, **k)
        return staticmethod(wrapped)

    saved = {}
    node_classes = [c for c in vars(_ops).values() if isinstance(c, type) and issubclass(c, torch.autograd.Function)
                    and c is not torch.autograd.Function]
    for cls in node_classes:
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
            # kernels launched through the C-ABI have no ATen parent op, so they land in the second group with the model
            for title, d in (("ATen / library kernels launched INSIDE the engine's autograd nodes (what is left of PyTorch "
                              "on the adapter path)", inside),
                             ("everything else: the engine's own kernels (C-ABI launches) + the model's ops", outside)):
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
    kernel_ms = {}
    for k, v in agg.items():
        if "lyco::" in k or k.startswith("lyco"):
            short = k.replace("void ", "").replace("lyco::", "").split("(")[0]
            kernel_ms[short] = round(kernel_ms.get(short, 0.0) + v[1] / 1e3, 3)
    kernel_ms["_all_kernels_of_the_step"] = round(tot / 1e3, 2)
    if args.kernel_table:
        with open(args.kernel_table, "w") as fh:
            fh.write(f"one eager step, CUDA kernels by total device time (us); sum = {tot / 1e3:.2f} ms\n")
            for name, (cnt, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
                fh.write(f"{t / 1e3:10.3f} ms {100 * t / tot:5.1f}% {cnt:6d}  {name}\n")
    return gemm_kernel_ms, kernel_ms


# ------------------------------------------------------------------ reference on the GPU (PyTorch eager)
def gpu_eager_reference(args, unet, engine_net, static, engine_loss):
    """Time the UNMODIFIED reference (baseline/_ref, ``lycoris.kohya.create_network`` -> ``apply_to``) on the same
    model, the same adapter parameters (engine state_dict loaded into the reference network — the wire format is
    shared) and the same inputs, PyTorch-eager under autocast: the denominator of BASELINE.json's >=4x target."""
    import torch
    import torch.nn.functional as F

    wl = args.wl
    ref_kohya, notes = import_reference()
    if ref_kohya is None:
        return {"unavailable": notes}
    device = static["sample"].device
    sd = {k: v.detach().clone() for k, v in engine_net.state_dict().items()}
    engine_net.restore()
    try:
        net = create_network(ref_kohya, wl, unet)
        net.apply_to(None, unet, False, True)
        net.to(device)
        missing = net.load_state_dict(sd, strict=False)
        n_missing = len(missing.missing_keys) + len(missing.unexpected_keys)
        net.requires_grad_(True)
        net.train()

        def step():
            for p in net.parameters():
                p.grad = None
            with torch.autocast("cuda", dtype=torch.bfloat16):
                out = unet(static["sample"], static["timesteps"], static["context"], static.get("added_cond"))
            loss = F.mse_loss(out.float(), static["target"].float())
            loss.backward()
            return loss

        for _ in range(3):
            loss = step()
        torch.cuda.synchronize()
        ref_loss = float(loss)
        n = max(1, args.ref_steps)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n):
            step()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / n
        rel = abs(engine_loss - ref_loss) / max(abs(ref_loss), 1e-12)
        out = {
            "value": 1000.0 / ms, "unit": "steps/s", "ms_per_step": ms, "steps": n, "warmup": 3,
            "source": "unmodified reference from baseline/_ref (pip install --no-deps --target of /root/reference), "
                      "lycoris.kohya.create_network + apply_to, PyTorch eager, bf16 autocast, same model / parameters / inputs",
            "loss": ref_loss, "engine_loss": None if engine_loss != engine_loss else engine_loss,
            "loss_rel_diff": None if rel != rel else rel,
            "loss_check": "n/a" if engine_loss != engine_loss else ("ok" if rel <= 2e-2 else "FAILED"),
            "state_dict_key_mismatches": n_missing,
        }
        if notes and "ia3" in str(wl["kw"].get("preset", "")) + wl["note"]:
            out["reference_notes"] = notes
        for lora in list(getattr(net, "loras", [])):
            lora.restore()
        del net
        return out
    finally:
        torch.cuda.empty_cache()


def run_reference_gpu(args):
    """Standalone arm: only the unmodified reference on one GPU (for profiling it, or when the engine is absent)."""
    import torch

    import lycoris_b200.kohya as kohya
    from workloads.unet_skeleton import UNetSkeleton

    if int(os.environ.get("RANK", 0)) != 0:
        return
    wl = args.wl
    torch.cuda.set_device(0)
    device = torch.device("cuda", 0)
    cfg = model_cfg(wl["model"])
    torch.manual_seed(0)
    unet = UNetSkeleton(cfg).to(device=device, dtype=torch.bfloat16)
    unet.requires_grad_(False)
    unet.train()
    # the engine's network is built only to obtain the (seeded, perturbed) adapter parameters; it never runs
    net = create_network(kohya, wl, unet)
    net.apply_to(None, unet, False, True)
    net.to(device)
    perturb_zero_factors(net, 1)
    batch = unet.synthetic_batch(wl["batch"], "cpu", torch.bfloat16, seed=2, sample_size=args.sample_size or None)
    st = {k: (v.to(torch.bfloat16) if v.is_floating_point() else v).to(device) for k, v in batch.items()}
    for k in ("sample", "target"):
        st[k] = st[k].contiguous(memory_format=torch.channels_last)
    args.ref_steps = args.steps
    ref = gpu_eager_reference(args, unet, net, st, float("nan"))
    print(json.dumps({"metric": f"{cfg.name.upper()}-UNet+{wl['label']} fwd+bwd steps/sec", "impl": "reference-gpu",
                      "value": ref.get("value"), "unit": "steps/s", "ms_per_step": ref.get("ms_per_step"), "n_gpus": 1,
                      "steps": args.steps, "warmup": 3, "dtype": "bf16", "data": "synthetic", "detail": ref,
                      "config": {"workload": f"{wl['name']}: {cfg.name} + {wl['label']}, batch {wl['batch']}, eager ATen path"}}),
          flush=True)


# ------------------------------------------------------------------ reference on the host cores
def wrapped_layer_signatures(args):
    """Distinct (adapter class, base layer geometry, adapter settings, rows) of the workload with multiplicity,
    enumerated by building the engine's network on a META-device skeleton (no weights) and running a meta
    forward with hooks on the wrapped layers."""
    import torch
    import torch.nn as nn

    import lycoris_b200.kohya as kohya
    from workloads.unet_skeleton import UNetSkeleton

    wl = args.wl
    cfg = model_cfg(wl["model"])
    with torch.device("meta"):
        unet = UNetSkeleton(cfg)
    net = create_network(kohya, wl, unet)
    info = {}
    hooks = []
    for lora in net.loras:
        org = lora.org_module[0]

        def hook(m, inp, out, _l=lora):
            x = inp[0]
            kind = type(_l).__name__
            extra = ()
            if kind == "LokrModule":
                w1 = _l.lokr_w1 if _l.use_w1 else _l.lokr_w1_a
                extra = (int(w1.shape[0]), bool(_l.use_w2))
            elif kind == "IA3Module":
                extra = (bool(_l.train_input),)
            if isinstance(m, nn.Linear):
                geo = ("linear", m.out_features, m.in_features, 1, 1, 0, x.numel() // x.shape[-1])
            else:
                geo = ("conv", m.out_channels, m.in_channels, m.kernel_size[0], m.stride[0], m.padding[0], x.shape[-1])
            key = (kind, geo, int(getattr(_l, "lora_dim", 0) or 0), float(_l.alpha) if hasattr(_l, "alpha") else 0.0, extra)
            info[key] = info.get(key, 0) + 1

        hooks.append(org.register_forward_hook(hook))
    s = args.sample_size or cfg.sample_size
    b = wl["batch"]
    with torch.no_grad():
        unet(torch.zeros(b, cfg.in_channels, s, s, device="meta"), torch.zeros(b, device="meta", dtype=torch.long),
             torch.zeros(b, cfg.context_len, cfg.cross_attention_dim, device="meta"),
             torch.zeros(b, cfg.addition_embed_dim, device="meta") if cfg.addition_embed_dim else None)
    for h in hooks:
        h.remove()
    return info


def _reference_module_factory():
    """Adapter classes of the unmodified reference (baseline/_ref) when vendored — else the oracle port."""
    ref_kohya, _ = import_reference()
    if ref_kohya is None:
        return None
    import lycoris.modules.ia3 as r_ia3
    import lycoris.modules.locon as r_locon
    import lycoris.modules.loha as r_loha
    import lycoris.modules.lokr as r_lokr

    return {"LoConModule": r_locon.LoConModule, "LohaModule": r_loha.LohaModule, "LokrModule": r_lokr.LokrModule,
            "IA3Module": r_ia3.IA3Module}


def cpu_reference(args, budget_s=20.0, reps=3):
    """The reference's CPU path — its own adapter classes from baseline/_ref wrapped on fp32 base layers, forward +
    autograd backward — timed on this box's host cores on a BOUNDED sample of the workload: every distinct wrapped
    layer (class, geometry, settings) is run at TWO reduced row counts (best of ``reps`` each), a line
    t = a + b*rows is fitted — ``a`` is the M-independent weight-side work (factor products, kron, W + dW),
    ``b`` the per-row contraction cost — and extrapolated to the layer's full row count, times its multiplicity.
    Un-wrapped ops (attention, norms) are not counted, which flatters the CPU.  fp32 (the reference's CPU regime)."""
    import torch
    import torch.nn as nn

    # torchrun exports OMP_NUM_THREADS=1 to every rank; the CPU leg runs on rank 0 alone and is entitled to the host's
    # cores ("all the host threads it can use"): restore torch's own default of one thread per physical core
    try:
        avail = len(os.sched_getaffinity(0))
    except Exception:  # noqa: BLE001
        avail = os.cpu_count() or 1
    want = max(1, avail // 2)
    if torch.get_num_threads() < want and os.environ.get("OMP_NUM_THREADS", "") in ("", "1"):
        torch.set_num_threads(want)
    threads = torch.get_num_threads()
    classes = _reference_module_factory()
    kind = "reference" if classes is not None else "port"
    if classes is None:
        import lycoris_b200.modules as M  # constructor-compatible; only the shapes are used, the oracle does the math
        from oracle import lyco_oracle as O

        classes = {"LoConModule": M.LoConModule, "LohaModule": M.LohaModule, "LokrModule": M.LokrModule,
                   "IA3Module": M.IA3Module}
    sigs = wrapped_layer_signatures(args)
    batch = args.wl["batch"]
    t_start = time.time()
    total_full = 0.0
    weight_side = 0.0
    deadline = t_start + budget_s
    n_timed = 0
    for (cls_name, geo, dim, alpha, extra), count in sorted(sigs.items(), key=lambda kv: -kv[1]):
        gk, N, Kd, ks, stride, pad, m = geo
        torch.manual_seed(0)
        if gk == "linear":
            base = nn.Linear(Kd, N)
            rows_full = m
            rows = sorted({max(8, min(rows_full, 64)), max(8, min(rows_full, 256))})
            mk = [lambda r=r: torch.randn(r, Kd) for r in rows]
        else:
            base = nn.Conv2d(Kd, N, ks, stride, pad)
            side_full = m
            rows_full = batch * ((side_full + 2 * pad - ks) // stride + 1) ** 2
            sides = sorted({max(4, min(side_full, 8)), max(4, min(side_full, 16))})
            rows = [((sd_ + 2 * pad - ks) // stride + 1) ** 2 for sd_ in sides]
            mk = [lambda sd_=sd_: torch.randn(1, Kd, sd_, sd_) for sd_ in sides]
        base.requires_grad_(False)
        kw = {}
        if cls_name == "LokrModule":
            kw["factor"] = extra[0]
        if cls_name == "IA3Module":
            kw["train_on_input"] = extra[0]
        mod = classes[cls_name]("bench", base, 1.0, dim or 4, alpha or 1, 0.0, 0.0, 0.0, False, **kw)
        with torch.no_grad():
            for p in mod.parameters():
                if float(p.abs().sum()) == 0.0:
                    p.normal_(0, 0.01)
        mod.apply_to()
        ts = []
        over = time.time() > deadline
        for make in mk:
            best = None
            for _ in range(1 if over else reps):
                x = make().requires_grad_(True)
                for p in mod.parameters():
                    p.grad = None
                t0 = time.perf_counter()
                if kind == "reference":
                    y = base(x)
                else:
                    y = _oracle_forward(O, mod, base, x)
                y.float().pow(2).mean().backward()
                dt = time.perf_counter() - t0
                best = dt if best is None else min(best, dt)
            ts.append(best)
        if hasattr(mod, "restore"):
            mod.restore()
        if len(rows) == 2 and rows[1] > rows[0]:
            b_ = max(0.0, (ts[1] - ts[0]) / (rows[1] - rows[0]))
            a_ = max(0.0, ts[0] - b_ * rows[0])
        else:
            a_, b_ = 0.0, ts[0] / rows[0]
        total_full += (a_ + b_ * rows_full) * count
        weight_side += a_ * count
        n_timed += 1
    steps_per_s = 1.0 / total_full if total_full > 0 else 0.0
    return {
        "value": steps_per_s, "unit": "steps/s", "cores": threads, "kind": kind,
        "sample": f"{'unmodified reference modules (baseline/_ref)' if kind == 'reference' else 'oracle port'} on fp32 "
                  f"CPU base layers; {n_timed} distinct wrapped layers (class, geometry, settings) timed fwd+bwd at two "
                  f"row counts (<=64 / <=256 rows, <=8x8 / <=16x16 spatial), best of {reps}, line fit t = a + b*rows "
                  f"extrapolated to the full rows and multiplied by multiplicity; un-wrapped ops excluded; "
                  f"{time.time() - t_start:.1f}s of CPU work",
        "extrapolated_s_per_step": total_full,
        "weight_side_s_per_step": weight_side,
        "host_cpus": os.cpu_count(),
    }


def _oracle_forward(O, mod, base, x):
    conv = None
    if mod.module_type.startswith("conv"):
        conv = dict(stride=base.stride, padding=base.padding, dilation=base.dilation, groups=base.groups)
    algo = {"LoConModule": "locon", "LohaModule": "loha", "LokrModule": "lokr", "IA3Module": "ia3"}[type(mod).__name__]
    cfg = {"scale": getattr(mod, "scale", 1.0), "multiplier": 1.0}
    if algo == "ia3":
        cfg["train_on_input"] = mod.train_input
    return O.layer_forward(algo, x, base.weight, base.bias, dict(mod.named_parameters()), cfg, conv)


def run_reference_cpu(args):
    """--impl reference: each step is ONE bounded sample of the workload through the reference's CPU path (see
    cpu_reference); the line reports the median over the K timed samples."""
    rank = int(os.environ.get("RANK", 0))
    if rank != 0:
        return
    wl = args.wl
    cfg = model_cfg(wl["model"])
    per_step = max(2.0, min(args.cpu_seconds, 150.0 / max(1, args.steps + min(args.warmup, 1))))
    for _ in range(min(args.warmup, 1)):
        cpu_reference(args, budget_s=per_step, reps=1)
    vals = [cpu_reference(args, budget_s=per_step, reps=2) for _ in range(max(1, args.steps))]
    vals.sort(key=lambda r: r["value"])
    med = vals[len(vals) // 2]
    spread = (vals[-1]["value"] - vals[0]["value"]) / med["value"] if med["value"] else None
    result = {
        "metric": f"{cfg.name.upper()}-UNet+{wl['label']} fwd+bwd steps/sec", "value": med["value"], "unit": "steps/s",
        "n_gpus": int(os.environ.get("WORLD_SIZE", 1)), "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1000.0 * med["extrapolated_s_per_step"], "higher_is_better": True,
        "scaling": "strong" if wl.get("global_batch") else "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic", "impl": "reference",
        "config": {"workload": f"{wl['name']}: {cfg.name}-unet-skeleton + {wl['label']} ({wl['note']}), per-GPU batch "
                               f"{wl['batch']}, latents {args.sample_size or cfg.sample_size}^2, fwd+bwd, the reference's "
                               f"CPU path ({med['kind']}) on the host cores, fp32",
                   "global_batch": wl["batch"], "parallelism": "cpu", "sample": med["sample"],
                   "spread_over_steps": spread},
        "cpu_baseline": med,
        "e2e": {"value": med["value"], "unit": "steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(result), flush=True)


def main():
    import lycoris_b200.logging  # noqa: F401  (the package logger sets its level at import: import first, then quiet it)

    logging.getLogger("LyCORIS").setLevel(logging.ERROR)  # keep stdout to the ONE JSON line
    args = parse()
    if args.impl == "reference":
        run_reference_cpu(args)
    elif args.impl == "reference-gpu":
        run_reference_gpu(args)
    else:
        run_engine(args)


if __name__ == "__main__":
    main()
