"""GPU parity, per wrapped layer: the CUDA engine (through the public module API -> C-ABI) against
(a) the committed reference outputs in tests/golden and (b) the oracle run on the same device.

Tolerances (stated against the REFERENCE, whose own bf16 path rounds the base and delta
contractions separately and snaps dW onto W's bf16 grid):
  y, dx      : |err| <= 2^-6 * max|ref|   (two bf16 ulps at the tensor's scale) and mse < 5e-4,
               the bound the reference's own test/functional.py:12-16 uses for bf16;
  param grads: relative Frobenius error <= 3e-2 (sums of ~1e3 bf16-rounded products).
"""
import random

import pytest
import torch

from helpers import build_base, build_product_module, case_ids, load_cases, oracle_args, rel_err, seed_dylora

pytestmark = pytest.mark.gpu

Y_REL = 2.0 ** -6
G_REL = 3e-2


def _run_engine(case, regime):  # noqa: C901
    dev = "cuda"
    base = build_base(case, dev)
    mod = build_product_module(case, base).to(dev)
    if regime == "bf16":
        base.to(torch.bfloat16)
        mod.to(torch.bfloat16)
    elif regime == "autocast_bf16":
        base.to(torch.bfloat16)
    mod.apply_to()
    mod.train()
    x = case["x"].to(dev).clone().requires_grad_(True)
    seed_dylora(case)
    if regime == "autocast_bf16":
        with torch.autocast("cuda", dtype=torch.bfloat16):
            y = base(x)
    else:
        y = base(x)
    y.backward(case["dy"].to(dev).to(y.dtype))
    grads = {k: v.grad for k, v in mod.named_parameters() if v.grad is not None}
    mod.restore()
    return y.detach(), x.grad, grads


def _check(name, y, dx, grads, ref_y, ref_dx, ref_grads):
    for tag, a, b in (("y", y, ref_y), ("dx", dx, ref_dx)):
        a, b = a.float().cpu(), b.float().cpu()
        bound = Y_REL * float(b.abs().max())
        err = float((a - b).abs().max())
        assert err <= bound, (name, tag, err, bound)
        assert float(((a - b) ** 2).mean()) < 5e-4, (name, tag)
    assert set(grads) == set(ref_grads), (name, sorted(grads), sorted(ref_grads))
    for k, g in ref_grads.items():
        e = rel_err(grads[k].cpu(), g.cpu())
        assert e <= G_REL, (name, k, e)


@pytest.mark.parametrize("regime", ["bf16", "autocast_bf16"])
def test_engine_matches_reference_fixtures(regime):
    cases = load_cases(regime)
    worst = 0.0
    for name in case_ids(regime):
        case = cases[name]
        y, dx, grads = _run_engine(case, regime)
        assert y.dtype == case["y"].dtype
        _check(name, y, dx, grads, case["y"], case["dx"], case["grads"])
        worst = max(worst, float((y.float().cpu() - case["y"].float()).abs().max()))
    print(f"[{regime}] 32 cases, worst |y - y_ref| = {worst:.4g}")


@pytest.mark.parametrize("regime", ["bf16", "autocast_bf16"])
def test_engine_matches_oracle_on_device(regime):
    """Same comparison with the oracle evaluated on the GPU (cuBLAS/cuDNN eager, like the reference
    would run there) — isolates device-kernel differences from CPU-vs-GPU library differences."""
    from oracle import lyco_oracle as O

    cases = load_cases(regime)
    ac = torch.bfloat16 if regime == "autocast_bf16" else None
    for name in case_ids(regime):
        case = cases[name]
        algo, p, cfg, conv = oracle_args(case, "cuda")
        oy, odx, og = O.layer_forward_backward(algo, case["x"].cuda(), case["weight"].cuda(), case["bias"].cuda(),
                                               p, cfg, case["dy"].cuda(), conv, ac)
        ref_grads = {}
        for k, g in og.items():
            if isinstance(g, list):
                for i, t in enumerate(g):
                    if t is not None:
                        ref_grads[f"{k}.{i}"] = t
            elif g is not None:
                ref_grads[k] = g
        y, dx, grads = _run_engine(case, regime)
        _check(name, y, dx, grads, oy, odx, ref_grads)


def test_fp32_regime_linear_cases_match_reference_fixtures():
    """fp32 base + fp32 adapter, no autocast (cfg #1's regime): Linear layers contract as three bf16 products
    accumulated in fp32 on the engine — checked against the reference's fp32 CPU outputs at 1e-4."""
    from lycoris_b200.engine import _lib

    cases = load_cases("fp32")
    for name in case_ids("fp32"):
        if not name.endswith("/linear"):
            continue
        case = cases[name]
        before = _lib.launch_count()
        y, dx, grads = _run_engine(case, "fp32")
        assert _lib.launch_count() >= before + 9, name  # 3 contractions x 3 bf16 products
        for tag, a_, b_ in (("y", y, case["y"]), ("dx", dx, case["dx"])):
            err = float((a_.float().cpu() - b_.float()).abs().max())
            assert err <= 1e-4 * max(1.0, float(b_.abs().max())), (name, tag, err)
        for k_, g in case["grads"].items():
            assert rel_err(grads[k_].cpu(), g) <= 1e-3, (name, k_, rel_err(grads[k_].cpu(), g))


def test_cfg1_locon_linear768_fp32_on_gpu():
    """BASELINE.json configs[0] in its own regime (fp32, no autocast) through the generic wrapper on the GPU."""
    import torch.nn as nn

    from helpers import load_cfg1
    from lycoris_b200.wrapper import LycorisNetwork, create_lycoris

    c = load_cfg1()
    base = nn.Sequential(nn.Linear(768, 768))
    base[0].weight.data = c["weight"].clone()
    base[0].bias.data = c["bias"].clone()
    base.cuda().requires_grad_(False)
    LycorisNetwork.apply_preset({"target_module": ["Linear"], "target_name": []})
    net = create_lycoris(base, 1.0, linear_dim=4, linear_alpha=1, algo="locon")
    net.apply_to()
    net.cuda()
    with torch.no_grad():
        for k_, v in net.loras[0].named_parameters():
            v.copy_(c["params"][k_])
    y = base(c["x"].cuda())
    loss = y.float().pow(2).mean()
    loss.backward()
    net.restore()
    assert float((y[0].detach().cpu() - c["y0"]).abs().max()) <= 1e-4 * float(c["y0"].abs().max())
    assert abs(float(loss) - float(c["loss"])) <= 1e-5 * abs(float(c["loss"]))
    for k_, g in c["grads"].items():
        mine = dict(net.loras[0].named_parameters())[k_].grad.cpu()
        assert rel_err(mine, g) <= 1e-3, (k_, rel_err(mine, g))


def test_fp32_base_under_autocast_runs():
    """fp32 base weights + fp32 adapter under torch.autocast(bf16): operands are cast to bf16 like
    autocast would do for F.linear; compared with the oracle under the same autocast."""
    from oracle import lyco_oracle as O

    cases = load_cases("fp32")
    for name in ("locon/linear", "lokr_full/linear", "loha/conv3", "ia3_out/linear"):
        case = cases[name]
        algo, p, cfg, conv = oracle_args(case, "cuda")
        oy, odx, og = O.layer_forward_backward(algo, case["x"].cuda(), case["weight"].cuda(), case["bias"].cuda(),
                                               p, cfg, case["dy"].cuda(), conv, torch.bfloat16)
        base = build_base(case, "cuda")
        mod = build_product_module(case, base).cuda()
        mod.apply_to()
        x = case["x"].cuda().requires_grad_(True)
        with torch.autocast("cuda", dtype=torch.bfloat16):
            y = base(x)
        y.backward(case["dy"].cuda().to(y.dtype))
        mod.restore()
        assert float((y.detach().float() - oy.float()).abs().max()) <= Y_REL * float(oy.float().abs().max()), name
        assert float((x.grad.float() - odx.float()).abs().max()) <= Y_REL * float(odx.float().abs().max()) + 1e-6, name


def test_stacked_wrappers_are_additive():
    """test/wrapper.py:233-287 of the reference: stacked == base + delta1 + delta2."""
    import torch.nn as nn

    import lycoris_b200 as L

    torch.manual_seed(0)
    base = nn.Linear(64, 96).cuda().to(torch.bfloat16)
    for p in base.parameters():
        p.requires_grad_(False)
    x = torch.randn(4, 7, 64, device="cuda", dtype=torch.bfloat16)
    y0 = base(x)
    a = L.LoConModule("a", base, 1.0, 4, 2).cuda().to(torch.bfloat16)
    b = L.LokrModule("b", base, 1.0, 100000, 1, factor=4).cuda().to(torch.bfloat16)
    with torch.no_grad():
        a.lora_up.weight.normal_(0, 0.05)
        b.lokr_w2.normal_(0, 0.05)
    a.apply_to()
    ya = base(x)
    a.restore()
    b.apply_to()
    yb = base(x)
    b.restore()
    a.apply_to()
    b.apply_to()
    yab = base(x)
    b.restore()
    a.restore()
    expect = y0.float() + (ya.float() - y0.float()) + (yb.float() - y0.float())
    assert float((yab.float() - expect).abs().max()) <= 2 * Y_REL * float(expect.abs().max())
    assert torch.equal(base(x), y0)


def test_zero_init_adapter_is_identity_and_engine_was_used():
    import torch.nn as nn

    import lycoris_b200 as L
    from lycoris_b200.engine import _lib

    base = nn.Linear(128, 128).cuda().to(torch.bfloat16)
    x = torch.randn(32, 128, device="cuda", dtype=torch.bfloat16)
    y0 = base(x)
    before = _lib.launch_count()
    m = L.LoConModule("a", base, 1.0, 8, 4).cuda().to(torch.bfloat16)
    m.apply_to()
    y1 = base(x)
    m.restore()
    assert _lib.launch_count() >= before + 2, "the CUDA extension did not run"
    # dW == 0 -> W' == W exactly; single fp32-accumulated contraction vs cuBLAS: <= 1 bf16 ulp apart
    assert float((y1.float() - y0.float()).abs().max()) <= 2.0 ** -7 * float(y0.float().abs().max())


CONV_ENGINE_CASES = [
    # (algo ctor, Nb, C, O, H, W, k, stride)
    ("lokr", 2, 64, 128, 12, 10, 3, 1),
    ("lokr", 2, 128, 64, 9, 9, 3, 1),
    ("locon", 3, 64, 64, 8, 8, 3, 1),
    ("loha", 2, 64, 72, 8, 8, 3, 1),     # O not a multiple of 64: dgrad takes the library path
    ("lokr", 2, 64, 64, 16, 16, 3, 2),   # stride 2: fprop + wgrad on the engine, dgrad library
]


@pytest.mark.parametrize("layout", ["nhwc", "nchw", "nchw_fp32"])
@pytest.mark.parametrize("algo,Nb,C,O,H,W,k,stride", CONV_ENGINE_CASES)
def test_conv_implicit_gemm_layer_matches_oracle(algo, Nb, C, O, H, W, k, stride, layout):
    """3x3 convolutions on channels_last activations run the TMA-im2col implicit GEMM (fprop, dgrad as
    fprop on dY with the flipped filter, wgrad) — compared with the oracle (cuDNN eager) on device."""
    import torch.nn as nn

    import lycoris_b200 as L
    from lycoris_b200.engine import _lib
    from oracle import lyco_oracle as O_

    torch.manual_seed(C + O + H)
    base = nn.Conv2d(C, O, k, stride, k // 2).cuda().to(torch.bfloat16)
    for p in base.parameters():
        p.requires_grad_(False)
    if algo == "lokr":
        mod = L.LokrModule("c", base, 1.0, 100000, 1, factor=8).cuda()
        with torch.no_grad():
            mod.lokr_w2.normal_(0, 0.02)
        oalgo, cfg = "lokr", {"scale": mod.scale, "multiplier": 1.0}
    elif algo == "locon":
        mod = L.LoConModule("c", base, 1.0, 8, 4).cuda()
        with torch.no_grad():
            mod.lora_up.weight.normal_(0, 0.05)
        oalgo, cfg = "locon", {"scale": mod.scale, "multiplier": 1.0}
    else:
        mod = L.LohaModule("c", base, 1.0, 4, 2).cuda()
        with torch.no_grad():
            mod.hada_w2_a.normal_(0, 0.1)
        oalgo, cfg = "loha", {"scale": mod.scale, "multiplier": 1.0}
    # nhwc: activations already channels_last; nchw: PyTorch's default layout (one engine transpose pass in, the
    # epilogue writes NCHW out); nchw_fp32: fp32 input under autocast, cast fused into that transpose pass
    x = torch.randn(Nb, C, H, W, device="cuda", dtype=torch.float32 if layout == "nchw_fp32" else torch.bfloat16)
    if layout == "nhwc":
        x = x.contiguous(memory_format=torch.channels_last)
    x.requires_grad_(True)
    before = _lib.launch_count()
    mod.apply_to()
    with torch.autocast("cuda", dtype=torch.bfloat16):
        y = base(x)
    dy = torch.randn_like(y)
    y.backward(dy)
    mod.restore()
    engine_dgrad = stride == 1 and O % 64 == 0
    nhwc = lambda t: t.is_contiguous(memory_format=torch.channels_last)  # noqa: E731
    layout_passes = int(not nhwc(x)) + int(not nhwc(dy))  # one engine transpose(+cast) pass per NCHW operand
    # merge + fprop + (dgrad) + wgrad + factor grads (+ layout passes)
    assert _lib.launch_count() - before >= (5 if engine_dgrad else 4) + layout_passes
    assert y.dtype == torch.bfloat16 and x.grad.dtype == x.dtype and x.grad.shape == x.shape
    if layout != "nhwc" and (y.shape[2] * y.shape[3]) % 32 == 0:
        assert y.is_contiguous(), "NCHW in -> NCHW out (written by the epilogue)"
        if engine_dgrad:
            assert x.grad.is_contiguous()
    p = {kk: v.detach() for kk, v in mod.named_parameters()}
    conv = dict(stride=base.stride, padding=base.padding, dilation=base.dilation, groups=1)
    oy, odx, og = O_.layer_forward_backward(oalgo, x.detach().to(torch.bfloat16), base.weight, base.bias, p, cfg, dy,
                                            conv, torch.bfloat16)
    assert float((y.detach().float() - oy.float()).abs().max()) <= Y_REL * float(oy.float().abs().max())
    assert float((x.grad.float() - odx.float()).abs().max()) <= Y_REL * float(odx.float().abs().max())
    for kk, g in og.items():
        assert rel_err(getattr(mod, kk.split(".")[0]).weight.grad if "." in kk else getattr(mod, kk).grad, g) <= G_REL, kk


LOWRANK_TC_CASES = [
    ("locon", 16, "autocast"), ("locon", 16, "bf16"), ("locon", 64, "autocast"),
    ("loha", 32, "autocast"), ("loha", 8, "bf16"), ("dylora", 16, "autocast"),
]


@pytest.mark.parametrize("algo,r,regime", LOWRANK_TC_CASES)
def test_lowrank_tensor_core_path_matches_oracle(algo, r, regime):
    """r % 8 == 0 and a 16-bit product domain: the rank-r products and their gradients run as K = r / N = r
    contractions on the tcgen05 GEMM (RAW merge + grad_prep); compared with the oracle on device."""
    import torch.nn as nn

    import lycoris_b200 as L
    from lycoris_b200.engine import ops
    from oracle import lyco_oracle as O_

    assert ops._LOWRANK_TC
    torch.manual_seed(r)
    base = nn.Linear(640, 1280).cuda().to(torch.bfloat16)
    for p in base.parameters():
        p.requires_grad_(False)
    if algo == "locon":
        mod = L.LoConModule("t", base, 1.0, r, r / 2).cuda()
        with torch.no_grad():
            mod.lora_up.weight.normal_(0, 0.05)
        cfg = {"scale": mod.scale, "multiplier": 1.0}
        params = lambda: {"lora_up.weight": mod.lora_up.weight, "lora_down.weight": mod.lora_down.weight}  # noqa: E731
    elif algo == "loha":
        mod = L.LohaModule("t", base, 1.0, r, r / 2).cuda()
        with torch.no_grad():
            mod.hada_w2_a.normal_(0, 0.1)
        cfg = {"scale": mod.scale, "multiplier": 1.0}
        params = lambda: {k: getattr(mod, k) for k in ("hada_w1_a", "hada_w1_b", "hada_w2_a", "hada_w2_b")}  # noqa: E731
    else:
        mod = L.DyLoraModule("t", base, 1.0, r, r / 2, block_size=8).cuda()
        with torch.no_grad():
            for u in mod.up_list:
                u.normal_(0, 0.05)
        cfg = {"alpha": mod.alpha, "multiplier": 1.0, "scale": 1.0}
        params = lambda: {"up_list": list(mod.up_list), "down_list": list(mod.down_list)}  # noqa: E731
    ac = None
    if regime == "bf16":
        mod.to(torch.bfloat16)
    else:
        ac = torch.bfloat16
    x = torch.randn(4, 77, 640, device="cuda", dtype=torch.bfloat16, requires_grad=True)
    dy = torch.randn(4, 77, 1280, device="cuda", dtype=torch.bfloat16)
    mod.apply_to()
    random.seed(5)
    if ac is not None:
        with torch.autocast("cuda", dtype=ac):
            y = base(x)
    else:
        y = base(x)
    y.backward(dy)
    mod.restore()
    if algo == "dylora":
        random.seed(5)
        cfg["b"] = O_.draw_dylora_block(mod.block_count)
    oy, odx, og = O_.layer_forward_backward("dylora" if algo == "dylora" else algo, x.detach(), base.weight,
                                            base.bias, {k: (v if isinstance(v, list) else v.detach()) for k, v in params().items()},
                                            cfg, dy, None, ac)
    assert float((y.detach().float() - oy.float()).abs().max()) <= Y_REL * float(oy.float().abs().max())
    assert float((x.grad.float() - odx.float()).abs().max()) <= Y_REL * float(odx.float().abs().max())
    for k, g in og.items():
        if isinstance(g, list):
            mine = [p.grad for p in getattr(mod, k)]
            for a, b in zip(mine, g):
                assert (a is None) == (b is None), k
                if b is not None:
                    assert rel_err(a, b) <= G_REL, k
        else:
            tgt = mod.lora_up.weight if k == "lora_up.weight" else mod.lora_down.weight if k == "lora_down.weight" else getattr(mod, k)
            assert rel_err(tgt.grad, g) <= G_REL, (k, rel_err(tgt.grad, g))


def test_cfg1_locon_linear768_on_gpu_under_autocast():
    """configs[0] through the generic wrapper on the GPU: fp32 base + fp32 adapter under autocast(bf16)
    against the reference's fp32 CPU result, at bf16 tolerance (the engine has no fp32-operand GEMM)."""
    import torch.nn as nn

    from helpers import load_cfg1
    from lycoris_b200.wrapper import LycorisNetwork, create_lycoris

    c = load_cfg1()
    base = nn.Sequential(nn.Linear(768, 768))
    base[0].weight.data = c["weight"].clone()
    base[0].bias.data = c["bias"].clone()
    base.cuda().requires_grad_(False)
    LycorisNetwork.apply_preset({"target_module": ["Linear"], "target_name": []})
    net = create_lycoris(base, 1.0, linear_dim=4, linear_alpha=1, algo="locon")
    net.apply_to()
    net.cuda()
    with torch.no_grad():
        for k, v in net.loras[0].named_parameters():
            v.copy_(c["params"][k])
    with torch.autocast("cuda", dtype=torch.bfloat16):
        y = base(c["x"].cuda())
    loss = y.float().pow(2).mean()
    loss.backward()
    net.restore()
    ref_y = c["y0"]
    assert float((y[0].detach().float().cpu() - ref_y).abs().max()) <= 2.0 ** -6 * float(ref_y.abs().max())
    assert abs(float(loss) - float(c["loss"])) <= 1e-2 * abs(float(c["loss"]))
    for k, g in c["grads"].items():
        mine = dict(net.loras[0].named_parameters())[k].grad.cpu()
        assert rel_err(mine, g) <= 5e-2, (k, rel_err(mine, g))


@pytest.mark.parametrize("regime", ["bf16", "autocast_bf16"])
def test_engine_option_variants_match_reference_fixtures(regime):
    """Trainable scalar, DoRA on either axis, Tucker conv, multiplier 0.5: W' is assembled by PyTorch ops
    following the reference's sequence and the three contractions run on the engine; outputs and every
    parameter gradient (incl. scalar / dora_scale / lora_mid) against the reference fixtures."""
    from lycoris_b200.engine import _lib

    cases = load_cases(regime, "options")
    for name in case_ids(regime, "options"):
        case = cases[name]
        before = _lib.launch_count()
        y, dx, grads = _run_engine(case, regime)
        if "/linear" in name:  # the 8-channel fixture conv has no TMA-legal layout; its contractions are cuDNN's
            assert _lib.launch_count() > before, name
        _check(name, y, dx, grads, case["y"], case["dx"], case["grads"])


@pytest.mark.parametrize("regime", ["bf16", "autocast_bf16"])
def test_engine_tucker_and_64_channel_variants_match_reference_fixtures(regime):
    """Round-2 fixtures: LoHa-Tucker, LoKr-Tucker, LoCon-Tucker, DoRA (lyco_dora_fwd / lyco_dora_bwd) and plain LoKr.
    The conv3c64 cases have 64 input channels, so fprop / dgrad / wgrad run on conv_sm100_kernel — an option-variant
    layer on the engine's own convolution kernels, against outputs of the reference itself."""
    from lycoris_b200.engine import _lib

    cases = load_cases(regime, "tucker")
    for name in case_ids(regime, "tucker"):
        case = cases[name]
        before = _lib.launch_count()
        y, dx, grads = _run_engine(case, regime)
        if name.endswith("conv3c64"):
            assert _lib.launch_count() >= before + 5, name  # relayouts + fprop + dgrad + wgrad at least
        _check(name, y, dx, grads, case["y"], case["dx"], case["grads"])


def test_conv_engine_with_channel_sliced_gradient_and_input():
    """Non-dense NHWC-looking tensors (channel slices of a concatenation — what torch.cat's backward hands to the
    upsampler convolution of a UNet) must be densified, not passed through: regression test for a layout pass that
    trusted Tensor.to(memory_format=channels_last) to copy."""
    import torch.nn as nn

    import lycoris_b200 as L

    torch.manual_seed(3)
    base = nn.Conv2d(64, 64, 3, 1, 1).cuda().to(torch.bfloat16)
    for p in base.parameters():
        p.requires_grad_(False)
    mod = L.LokrModule("c", base, 1.0, 100000, 1, factor=8).cuda()
    with torch.no_grad():
        mod.lokr_w2.normal_(0, 0.02)
    big = torch.randn(2, 96, 16, 16, device="cuda").contiguous(memory_format=torch.channels_last)
    big.requires_grad_(True)
    skip = torch.randn(2, 32, 16, 16, device="cuda", dtype=torch.bfloat16).contiguous(memory_format=torch.channels_last)

    def run():
        big.grad = None
        with torch.autocast("cuda", dtype=torch.bfloat16):
            y = base(big[:, 16:80])                 # non-dense input: channel slice of an NHWC tensor
            z = torch.cat([y, skip], dim=1)         # its gradient comes back as a channel slice, too
        (z.float() * torch.linspace(-1, 1, z.numel(), device="cuda").view_as(z)).sum().backward()
        return y.detach().float(), big.grad.detach().clone(), None if mod.lokr_w2.grad is None else mod.lokr_w2.grad.clone()

    ref_w = (base.weight.float() + mod.get_weight(base.weight.shape).float() * mod.multiplier).to(torch.bfloat16)
    plain = nn.Conv2d(64, 64, 3, 1, 1).cuda().to(torch.bfloat16)
    with torch.no_grad():
        plain.weight.copy_(ref_w)
        plain.bias.copy_(base.bias)
    mod.apply_to()
    y, dx, gw = run()
    mod.restore()
    big.grad = None
    with torch.autocast("cuda", dtype=torch.bfloat16):
        y2 = plain(big[:, 16:80])
        z2 = torch.cat([y2, skip], dim=1)
    (z2.float() * torch.linspace(-1, 1, z2.numel(), device="cuda").view_as(z2)).sum().backward()
    assert rel_err(y, y2.detach().float()) <= 1e-2
    assert rel_err(dx, big.grad) <= 2e-2, rel_err(dx, big.grad)
    assert gw is not None and float(gw.abs().sum()) > 0
