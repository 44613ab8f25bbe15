"""GPU parity at the HEADLINE sizes (BASELINE.json cfg #3 / #4 / #5 shapes), engine vs the oracle on the same device.

Round-1 parity ran on the 32/64-channel toy model and on golden fixtures whose convolutions have 16 / 32 input
channels (cuDNN path).  Here:
  * single wrapped layers at SDXL size — Linear 1280->1280 / 1280->10240 / 5120->1280 at M = 8192, Linear 2048->1280
    at M = 616, Conv3x3 1280->1280 @32^2 b8 and 320->320 @128^2 b8 (M = 131072), conv_shortcut 1x1 — LoKr f8
    full-dim, LoHa d32, LoCon d16, IA3, with the per-layer tolerances of test_gpu_layers.py (Y_REL = 2^-6 * max|ref|,
    G_REL = 3e-2) — every one of them runs on conv_sm100_kernel / gemm_sm100_kernel;
  * one BasicTransformerBlock d = 1280 at M = 8192 and one ResnetBlock2D 1280 -> 1280 @32^2 batch 8, whole-block
    fwd+bwd, LoKr f8 and LoHa d32 (relative Frobenius bounds, stated below);
  * cfg #5: locon + loha + lokr + ia3 through kohya.create_network with the bench's own preset file.
"""
import json
import os
import random
import statistics

import pytest
import torch
import torch.nn as nn
import torch.nn.functional as F

from helpers import oracle_patch_network, rel_err

pytestmark = pytest.mark.gpu

Y_REL = 2.0 ** -6
G_REL = 3e-2
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _log(name, rec):
    """Measured errors go to gpurun_out/ (scratch) so the bounds asserted here can be tightened to what they measure."""
    d = os.path.join(ROOT, "gpurun_out")
    if os.path.isdir(d):
        with open(os.path.join(d, "headline_parity_errors.jsonl"), "a") as fh:
            fh.write(json.dumps({"case": name, **rec}) + "\n")


def _mk_layer(kind, N, K, k=1, stride=1, pad=0):
    torch.manual_seed(0)
    base = nn.Linear(K, N) if kind == "linear" else nn.Conv2d(K, N, k, stride, pad)
    base = base.cuda().to(torch.bfloat16)
    base.requires_grad_(False)
    return base


def _mk_adapter(algo, base):
    import lycoris_b200 as L

    torch.manual_seed(1)
    if algo == "lokr":
        mod = L.LokrModule("t", base, 1.0, 100000, 1, factor=8)
    elif algo == "loha":
        mod = L.LohaModule("t", base, 1.0, 32, 16)
    elif algo == "locon":
        mod = L.LoConModule("t", base, 1.0, 16, 8)
    elif algo == "ia3":
        from lycoris_b200.modules.ia3 import IA3Module

        mod = IA3Module("t", base, 1.0, train_on_input=False)
    mod = mod.cuda()
    g = torch.Generator().manual_seed(3)
    with torch.no_grad():
        for p in mod.parameters():
            if float(p.abs().sum()) == 0.0:
                p.copy_((torch.randn(p.shape, generator=g) * 0.02).to(p))
    return mod


def _oracle_call(algo, mod, base, x, dy):
    from oracle import lyco_oracle as O

    conv = None
    if isinstance(base, nn.Conv2d):
        conv = dict(stride=base.stride, padding=base.padding, dilation=base.dilation, groups=base.groups)
    cfg = {"scale": getattr(mod, "scale", 1.0), "multiplier": 1.0}
    if algo == "ia3":
        cfg["train_on_input"] = mod.train_input
    p = {k: v.detach() for k, v in mod.named_parameters()}
    return O.layer_forward_backward(algo, x, base.weight, base.bias, p, cfg, dy, conv, torch.bfloat16)


LAYERS = [
    # name, kind, N, K, k, stride, pad, input shape
    ("attn_1280", "linear", 1280, 1280, 1, 1, 0, (8, 1024, 1280)),
    ("geglu_1280", "linear", 10240, 1280, 1, 1, 0, (8, 1024, 1280)),
    ("ffout_1280", "linear", 1280, 5120, 1, 1, 0, (8, 1024, 5120)),
    ("xattn_kv", "linear", 1280, 2048, 1, 1, 0, (8, 77, 2048)),
    ("attn_640", "linear", 640, 640, 1, 1, 0, (8, 4096, 640)),
    ("res_1280_32", "conv", 1280, 1280, 3, 1, 1, (8, 1280, 32, 32)),
    ("res_320_128", "conv", 320, 320, 3, 1, 1, (8, 320, 128, 128)),
    ("shortcut_1x1", "conv", 640, 1280, 1, 1, 0, (8, 1280, 64, 64)),
    ("down_s2_640", "conv", 640, 640, 3, 2, 1, (8, 640, 64, 64)),
]
ALGOS = ["lokr", "loha", "locon", "ia3"]
# LoKr runs every shape; the other algorithms run a Linear, the big GEGLU projection and one convolution
MATRIX = [(a, l) for a in ALGOS for l in LAYERS
          if a == "lokr" or l[0] in ("attn_1280", "geglu_1280", "res_1280_32", "xattn_kv")]


@pytest.mark.parametrize("algo,layer", MATRIX, ids=[f"{a}-{l[0]}" for a, l in MATRIX])
def test_sdxl_size_layer_matches_oracle(algo, layer):
    from lycoris_b200.engine import _lib

    name, kind, N, K, k, stride, pad, xshape = layer
    base = _mk_layer(kind, N, K, k, stride, pad)
    mod = _mk_adapter(algo, base)
    g = torch.Generator().manual_seed(7)
    x = torch.randn(xshape, generator=g).cuda().to(torch.bfloat16)
    if kind == "conv":
        x = x.contiguous(memory_format=torch.channels_last)
    with torch.no_grad():
        yshape = base(x[:1]).shape
    dy = (torch.randn((xshape[0], *yshape[1:]), generator=g) * 0.1).cuda().to(torch.bfloat16)
    oy, odx, og = _oracle_call(algo, mod, base, x, dy)

    before = _lib.launch_count()
    mod.apply_to()
    xe = x.clone().requires_grad_(True)
    with torch.autocast("cuda", dtype=torch.bfloat16):
        y = base(xe)
    y.backward(dy)
    mod.restore()
    assert _lib.launch_count() >= before + 4, "engine kernels did not run"

    rec = {}
    for tag, a_, b_ in (("y", y.detach(), oy), ("dx", xe.grad, odx)):
        a_, b_ = a_.float(), b_.float()
        err, bound = float((a_ - b_).abs().max()), Y_REL * float(b_.abs().max())
        rec[tag] = err / max(float(b_.abs().max()), 1e-30)
        assert err <= bound, (name, algo, tag, err, bound)
    for kname, ref in og.items():
        mine = dict(mod.named_parameters())[kname].grad
        e = rel_err(mine, ref)
        rec["g_" + kname] = e
        assert e <= G_REL, (name, algo, kname, e)
    _log(f"layer/{algo}/{name}", rec)


def _block_case(which):
    from workloads.unet_skeleton import BasicTransformerBlock, ResnetBlock2D

    torch.manual_seed(0)
    if which == "transformer":
        blk = BasicTransformerBlock(1280, 20, 2048)
        g = torch.Generator().manual_seed(5)
        inputs = (torch.randn(8, 1024, 1280, generator=g), torch.randn(8, 77, 2048, generator=g))
    else:
        blk = ResnetBlock2D(1280, 1280, 1280, 32)
        g = torch.Generator().manual_seed(5)
        inputs = (torch.randn(8, 1280, 32, 32, generator=g), torch.randn(8, 1280, generator=g))
    blk = blk.cuda().to(torch.bfloat16)
    blk.requires_grad_(False)
    blk.train()
    inputs = tuple(t.cuda().to(torch.bfloat16) for t in inputs)
    if which == "resnet":
        inputs = (inputs[0].contiguous(memory_format=torch.channels_last), inputs[1])
    return blk, inputs


@pytest.mark.parametrize("which", ["transformer", "resnet"])
@pytest.mark.parametrize("algo", ["lokr", "loha"])
def test_sdxl_size_block_matches_oracle_network(which, algo):
    """One BasicTransformerBlock (d = 1280, M = 8192, 10 wrapped Linear layers) / one ResnetBlock2D (1280 -> 1280 @32^2,
    batch 8: two 3x3 convolutions + time_emb_proj) wrapped through the generic wrapper, engine vs the same network
    with every wrapped layer routed through the oracle on the same device."""
    from lycoris_b200.wrapper import LycorisNetwork, create_lycoris

    blk, inputs = _block_case(which)
    LycorisNetwork.apply_preset({"target_module": ["Linear", "Conv2d"], "target_name": []})
    torch.manual_seed(1)
    if algo == "lokr":
        net = create_lycoris(blk, 1.0, linear_dim=100000, linear_alpha=1, algo="lokr", factor=8)
    else:
        net = create_lycoris(blk, 1.0, linear_dim=32, linear_alpha=16, conv_dim=16, conv_alpha=8, algo="loha")
    for lora in net.loras:
        net.add_module(lora.lora_name, lora)
    net.cuda()
    g = torch.Generator().manual_seed(3)
    with torch.no_grad():
        for p in net.parameters():
            if float(p.abs().sum()) == 0.0:
                p.copy_((torch.randn(p.shape, generator=g) * 0.02).to(p))
    net.requires_grad_(True)
    assert len(net.loras) == (10 if which == "transformer" else 3)

    def run():
        for p in net.parameters():
            p.grad = None
        x = inputs[0].clone().requires_grad_(True)
        with torch.autocast("cuda", dtype=torch.bfloat16):
            out = blk(x, inputs[1])
        loss = out.float().pow(2).mean()
        loss.backward()
        return float(loss), out.detach(), x.grad.detach(), {n: p.grad.detach().clone() for n, p in net.named_parameters()}

    undo = oracle_patch_network(net)
    ref_loss, ref_out, ref_dx, ref_g = run()
    undo()
    net.apply_to()
    loss, out, dx, grads = run()
    net.restore()

    errs = {k: rel_err(grads[k], ref_g[k]) for k in ref_g if float(ref_g[k].float().norm()) > 0}
    rec = {"loss_rel": abs(loss - ref_loss) / abs(ref_loss), "out": rel_err(out, ref_out), "dx": rel_err(dx, ref_dx),
           "g_median": statistics.median(errs.values()), "g_worst": max(errs.values())}
    _log(f"block/{which}/{algo}", rec)
    # measured on B200 (round 2): loss 2e-5, out 0.6-1.1e-2, dx 0.9-1.6e-2, gradient median 0.7-1.3e-2 — the bounds
    # leave 2x for run-to-run differences of the split-K atomics; an order of magnitude below the toy-network bounds
    assert rec["loss_rel"] <= 1e-3, rec
    assert rec["out"] <= 2e-2, rec
    assert rec["dx"] <= 3e-2, rec
    assert rec["g_median"] <= 3e-2, rec
    assert rec["g_worst"] <= 8e-2, (rec, max(errs.items(), key=lambda kv: kv[1]))


def test_cfg5_mixed_preset_network_matches_oracle_network():
    """BASELINE.json cfg #5 flavour through the kohya entry point with the bench's own preset file: locon on
    ResnetBlock2D / samplers, lokr f8 on FeedForward, loha on attention q/out + proj, **ia3** on to_k / to_v."""
    import bench
    import lycoris_b200.kohya as kohya
    from oracle.toy_models import ToyUNet

    wl = dict(bench.CONFIGS["cfg5"])
    torch.manual_seed(0)
    unet = ToyUNet().cuda().to(torch.bfloat16).to(memory_format=torch.channels_last)
    unet.requires_grad_(False)
    unet.train()
    net = bench.create_network(kohya, wl, unet)
    kinds = {type(l).__name__ for l in net.loras}
    assert kinds == {"LoConModule", "LohaModule", "LokrModule", "IA3Module"}, kinds
    for lora in net.loras:
        net.add_module(lora.lora_name, lora)
    net.cuda()
    g = torch.Generator().manual_seed(3)
    with torch.no_grad():
        for p in net.parameters():
            if float(p.abs().sum()) == 0.0:
                p.copy_((torch.randn(p.shape, generator=g) * 0.02).to(p))
    net.requires_grad_(True)
    batch = unet.synthetic_batch(2, "cpu", torch.bfloat16, seed=5)
    st = {k: (v.to(torch.bfloat16) if v.is_floating_point() else v).cuda() for k, v in batch.items()}
    st["sample"] = st["sample"].contiguous(memory_format=torch.channels_last)

    def run():
        for p in net.parameters():
            p.grad = None
        x = st["sample"].clone().requires_grad_(True)
        random.seed(11)
        with torch.autocast("cuda", dtype=torch.bfloat16):
            out = unet(x, st["timesteps"], st["context"])
        loss = F.mse_loss(out.float(), st["target"].float())
        loss.backward()
        return float(loss), out.detach(), x.grad.detach(), {n: p.grad.detach().clone() for n, p in net.named_parameters()
                                                               if p.grad is not None}

    undo = oracle_patch_network(net)
    ref_loss, ref_out, ref_dx, ref_g = run()
    undo()
    net.apply_to(None, unet, False, True)
    loss, out, dx, grads = run()
    net.restore()
    assert set(grads) == set(ref_g)
    errs = {k: rel_err(grads[k], ref_g[k]) for k in ref_g if float(ref_g[k].float().norm()) > 0}
    rec = {"loss_rel": abs(loss - ref_loss) / abs(ref_loss), "out": rel_err(out, ref_out), "dx": rel_err(dx, ref_dx),
           "g_median": statistics.median(errs.values()), "g_worst": max(errs.values())}
    _log("network/cfg5_toy", rec)
    # measured on B200: loss 7e-5, out 1.3e-2, dx 2.0e-2, gradient median 2.3e-2, worst 3.5e-2 (the round-1 bound on the
    # worst parameter gradient of the toy network was 0.25)
    assert rec["loss_rel"] <= 5e-3 and rec["out"] <= 3e-2 and rec["dx"] <= 4e-2, rec
    assert rec["g_median"] <= 5e-2, rec
    assert rec["g_worst"] <= 0.1, (rec, max(errs.items(), key=lambda kv: kv[1]))


# --------------------------------------------------------------------------- dropout on the GPU
def _locon_linear(rank_dropout=0.0, module_dropout=0.0, scale=False):
    import lycoris_b200 as L

    torch.manual_seed(0)
    base = nn.Linear(256, 128).cuda().to(torch.bfloat16)
    base.requires_grad_(False)
    mod = L.LoConModule("d", base, 1.0, 8, 4, 0.0, rank_dropout, module_dropout, rank_dropout_scale=scale).cuda().to(torch.bfloat16)
    with torch.no_grad():
        mod.lora_up.weight.normal_(0, 0.05)
    mod.train()
    return base, mod


@pytest.mark.parametrize("scale", [False, True])
def test_rank_dropout_locon_on_gpu_matches_reference_formula(scale):
    """Rebuild-mode rank dropout (locon.py:210-217): a Bernoulli mask over OUTPUT ROWS of dW drawn on the weight's
    device.  Same CUDA seed -> same mask; the expected result is built from the reference's formula."""
    base, mod = _locon_linear(rank_dropout=0.5, scale=scale)
    x = torch.randn(4, 32, 256, device="cuda", dtype=torch.bfloat16, requires_grad=True)
    dy = torch.randn(4, 32, 128, device="cuda", dtype=torch.bfloat16)
    mod.apply_to()
    torch.manual_seed(123)
    y = base(x)
    y.backward(dy)
    mod.restore()
    # reference formula with the same draw
    torch.manual_seed(123)
    up, down = mod.lora_up.weight.detach().clone().requires_grad_(True), mod.lora_down.weight.detach().clone().requires_grad_(True)
    w = up @ down
    drop = (torch.rand(w.size(0), device="cuda") > 0.5).to(w.dtype).view(-1, 1)
    assert 0 < float(drop.sum()) < w.size(0)
    if scale:
        drop = drop / drop.mean()
    dW = (w * drop) * mod.scalar.to(w) * mod.scale
    Wn = base.weight + dW
    xr = x.detach().clone().requires_grad_(True)
    yr = F.linear(xr, base.weight, base.bias) + F.linear(xr, Wn - base.weight)
    yr.backward(dy)
    assert float((y.float() - yr.float()).abs().max()) <= Y_REL * float(yr.float().abs().max())
    assert float((x.grad.float() - xr.grad.float()).abs().max()) <= Y_REL * float(xr.grad.float().abs().max())
    assert rel_err(mod.lora_up.weight.grad, up.grad) <= G_REL
    assert rel_err(mod.lora_down.weight.grad, down.grad) <= G_REL
    # dropped rows carry no adapter gradient at all
    dead = (drop.view(-1) == 0)
    assert float(mod.lora_up.weight.grad[dead].abs().max()) == 0.0


def test_module_dropout_on_gpu_is_a_per_call_coin():
    """module_dropout (locon.py:310-313): with probability p the call is the plain base layer."""
    base, mod = _locon_linear(module_dropout=0.5)
    x = torch.randn(8, 256, device="cuda", dtype=torch.bfloat16)
    plain = F.linear(x, base.weight, base.bias)
    mod.apply_to()
    torch.manual_seed(0)
    dropped = 0
    for _ in range(200):
        y = base(x)
        dropped += int(torch.equal(y, plain))
    mod.eval()
    assert not torch.equal(base(x), plain), "eval mode never drops"
    mod.restore()
    assert 70 <= dropped <= 130, dropped


@pytest.mark.parametrize("what", ["module_dropout", "dylora"])
def test_graph_capture_refuses_host_randomness(what):
    """A captured step would freeze the host coin / DyLoRA's random.randint into the graph: raise instead."""
    import lycoris_b200 as L

    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        if what == "module_dropout":
            base, mod = _locon_linear(module_dropout=0.3)
        else:
            torch.manual_seed(0)
            base = nn.Linear(256, 128).cuda().to(torch.bfloat16)
            base.requires_grad_(False)
            mod = L.DyLoraModule("d", base, 1.0, 8, 4, block_size=2).cuda().to(torch.bfloat16)
            mod.train()
        x = torch.randn(8, 256, device="cuda", dtype=torch.bfloat16)
        mod.apply_to()
        for _ in range(3):
            base(x)  # eager is fine
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with pytest.raises(RuntimeError, match="captured CUDA graph"):
            with torch.cuda.graph(g, stream=s):
                base(x)
        mod.restore()
    torch.cuda.synchronize()


# ---------------------------------------------------------------- text-encoder adapters on the engine (§8 f4)
def test_text_encoder_adapters_run_on_the_engine_and_match_oracle():
    """kohya.create_network with two CLIP-shaped text encoders (lora_te1_* / lora_te2_*): the adapters' Linear
    layers go through the same engine kernels as the UNet's; engine vs oracle-patched network on the GPU."""
    import lycoris_b200.kohya as kohya
    from lycoris_b200.engine import _lib
    from oracle.toy_models import ToyTextEncoder, ToyUNet

    torch.manual_seed(0)
    tes = [ToyTextEncoder(64, 2, 2).cuda().to(torch.bfloat16), ToyTextEncoder(128, 4, 1).cuda().to(torch.bfloat16)]
    unet = ToyUNet().cuda().to(torch.bfloat16)
    for m in (*tes, unet):
        m.requires_grad_(False)
    torch.manual_seed(1)
    net = kohya.create_network(1.0, 8, 4, None, tes, unet, algo="lokr", factor=4, preset="full")
    te_loras = net.text_encoder_loras
    assert te_loras and all(l.lora_name.startswith(("lora_te1_", "lora_te2_")) for l in te_loras)
    for lora in net.text_encoder_loras:
        net.add_module(lora.lora_name, lora)
    net.cuda()
    g = torch.Generator().manual_seed(3)
    with torch.no_grad():
        for p in net.parameters():
            if float(p.abs().sum()) == 0.0:
                p.copy_((torch.randn(p.shape, generator=g) * 0.02).to(p))
    net.requires_grad_(True)
    ids = torch.randint(0, 64, (4, 16), generator=g).cuda()

    def run():
        for p in net.parameters():
            p.grad = None
        with torch.autocast("cuda", dtype=torch.bfloat16):
            out = torch.cat([te(ids) for te in tes], dim=-1)
        loss = out.float().pow(2).mean()
        loss.backward()
        return float(loss), out.detach(), {n: p.grad.detach().clone() for n, p in net.named_parameters() if p.grad is not None}

    class _TE:  # oracle_patch_network walks `.loras`
        loras = te_loras

    undo = oracle_patch_network(_TE)
    ref_loss, ref_out, ref_g = run()
    undo()
    before = _lib.launch_count()
    net.apply_to(tes, unet, True, False)
    loss, out, grads = run()
    net.restore()
    assert _lib.launch_count() > before + len(te_loras), "text-encoder adapters did not reach the engine"
    assert abs(loss - ref_loss) <= 2e-2 * abs(ref_loss)
    assert rel_err(out, ref_out) <= 3e-2
    assert set(grads) == set(ref_g) and len(grads) >= 2 * len(te_loras)
    errs = [rel_err(grads[k], ref_g[k]) for k in ref_g if float(ref_g[k].float().norm()) > 0]
    assert statistics.median(errs) <= 5e-2 and max(errs) <= 0.25, (statistics.median(errs), max(errs))


# ------------------------------------------------------------- round-1 advisor findings, on the device
def test_foreign_patched_forward_is_not_dropped():
    """A base layer whose forward was already instance-patched (kohya networks.lora, an accelerate hook): the adapter
    must keep calling it (reference: always ``org_forward``) instead of contracting org.weight directly."""
    import lycoris_b200 as L

    torch.manual_seed(0)
    lin = nn.Linear(128, 128).cuda().to(torch.bfloat16)
    lin.requires_grad_(False)
    orig = lin.forward

    def foreign(x):
        return orig(x) + 1.0

    lin.forward = foreign
    mod = L.LoConModule("f", lin, 1.0, 8, 4).cuda().to(torch.bfloat16)
    mod.apply_to()
    assert not mod._is_outermost_on_plain_forward()
    x = torch.randn(16, 128, device="cuda", dtype=torch.bfloat16)
    y = lin(x)
    mod.restore()
    # lora_up is zero-initialised: the adapter adds nothing, the foreign +1 must still be there
    assert torch.allclose(y.float(), (F.linear(x, lin.weight, lin.bias) + 1.0).float(), atol=2e-2)


def test_ragged_out_features_and_empty_batch():
    """N % 8 != 0 (Linear(768, 10)) has no TMA-addressable output row pitch -> library GEMM, not an error; an empty
    batch returns an empty result (the reference handles both)."""
    import lycoris_b200 as L

    torch.manual_seed(0)
    lin = nn.Linear(768, 10).cuda().to(torch.bfloat16)
    lin.requires_grad_(False)
    mod = L.LoConModule("r", lin, 1.0, 4, 2).cuda().to(torch.bfloat16)
    with torch.no_grad():
        mod.lora_up.weight.normal_(0, 0.05)
    mod.apply_to()
    x = torch.randn(32, 768, device="cuda", dtype=torch.bfloat16, requires_grad=True)
    y = lin(x)
    y.float().pow(2).mean().backward()
    W = lin.weight.float() + (mod.lora_up.weight.float() @ mod.lora_down.weight.float()) * mod.scale
    ref = F.linear(x.detach().float(), W, lin.bias.float())
    assert float((y.float() - ref).abs().max()) <= 2 * Y_REL * float(ref.abs().max())
    assert mod.lora_down.weight.grad is not None and x.grad is not None
    y0 = lin(torch.empty(0, 768, device="cuda", dtype=torch.bfloat16))
    assert y0.shape == (0, 10)
    mod.restore()


def test_trainable_base_layer_gets_its_gradient():
    """A base layer with requires_grad=True (joint fine-tune): its gradient comes from org_forward like in the
    reference — the engine must not swallow the base contraction."""
    import lycoris_b200 as L

    torch.manual_seed(0)
    lin = nn.Linear(128, 64).cuda().to(torch.bfloat16)
    mod = L.LoConModule("t", lin, 1.0, 8, 4).cuda().to(torch.bfloat16)
    with torch.no_grad():
        mod.lora_up.weight.normal_(0, 0.05)
    mod.apply_to()
    x = torch.randn(16, 128, device="cuda", dtype=torch.bfloat16)
    lin(x).float().pow(2).mean().backward()
    mod.restore()
    assert lin.weight.grad is not None and lin.bias.grad is not None and float(lin.weight.grad.abs().sum()) > 0
    assert mod.lora_up.weight.grad is not None
