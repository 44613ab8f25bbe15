"""Generate tests/golden/* from the REAL reference (run in the build container only:
``python oracle/gen_golden.py``; needs /root/reference, which does not exist on the GPU box).

For every case the unmodified reference module (lycoris.modules.*) is built on a seeded base
layer, its zero-initialised factors are perturbed so dW != 0, and forward + backward are run on
the CPU.  The same tensors go through ``oracle.lyco_oracle``; the script asserts the oracle
reproduces the reference (bit-exact here, same ATen calls), then stores inputs + reference outputs.
Also written: the factorization table, and structural fixtures (adapter names / classes / state
dict shapes produced by the reference wrapper + kohya adapter on a small UNet-shaped model).
"""

import json
import os
import random
import sys

import torch
import torch.nn as nn

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(1, "/root/reference")

import lycoris  # noqa: E402  (the reference)
import lycoris.kohya  # noqa: E402
from lycoris.functional.general import factorization as ref_factorization  # noqa: E402

from oracle import lyco_oracle as O  # noqa: E402
from oracle.toy_models import ToyUNet  # noqa: E402

import logging  # noqa: E402

logging.getLogger("LyCORIS").setLevel(logging.ERROR)
OUT = os.path.join(ROOT, "tests", "golden")

LAYERS = {
    "linear": dict(kind="linear", in_dim=64, out_dim=96, x=(2, 5, 64)),
    "conv3": dict(kind="conv", in_dim=16, out_dim=32, k=3, stride=1, pad=1, x=(2, 16, 8, 8)),
    "conv1": dict(kind="conv", in_dim=32, out_dim=16, k=1, stride=1, pad=0, x=(2, 32, 6, 6)),
    "conv3s2": dict(kind="conv", in_dim=16, out_dim=16, k=3, stride=2, pad=1, x=(2, 16, 8, 8)),
    # 64 input channels: TMA-im2col addressable, so the engine runs these on conv_sm100_kernel (round-2 fixtures,
    # written by oracle/gen_golden_tucker.py; not part of the layers_* / options_* loops below)
    "conv3c64": dict(kind="conv", in_dim=64, out_dim=64, k=3, stride=1, pad=1, x=(2, 64, 8, 8)),
}
ROUND1_LAYERS = ("linear", "conv3", "conv1", "conv3s2")
ALGOS = {
    "locon": dict(cls="LoConModule", dim=4, alpha=2.0, kw={}),
    "loha": dict(cls="LohaModule", dim=4, alpha=2.0, kw={}),
    "lokr_full": dict(cls="LokrModule", dim=100000, alpha=1.0, kw={"factor": 4}),
    "lokr_lowrank": dict(cls="LokrModule", dim=2, alpha=1.0, kw={"factor": 4}),
    "lokr_both": dict(cls="LokrModule", dim=1, alpha=1.0, kw={"factor": 8, "decompose_both": True}),
    "ia3_out": dict(cls="IA3Module", dim=4, alpha=1.0, kw={"train_on_input": False}),
    "ia3_in": dict(cls="IA3Module", dim=4, alpha=1.0, kw={"train_on_input": True}),
    "dylora": dict(cls="DyLoraModule", dim=8, alpha=4.0, kw={"block_size": 2}),
}
# option variants: trainable scalar, DoRA on either axis, Tucker conv,
# non-unit multiplier — served by the engine's PyTorch-assembled-W' path + the tcgen05 contractions
OPTION_CASES = {
    "locon_scalar/linear": dict(algo="locon", layer="linear", kw={"use_scalar": True}),
    "locon_scalar/conv3": dict(algo="locon", layer="conv3", kw={"use_scalar": True}),
    "locon_dora/linear": dict(algo="locon", layer="linear", kw={"weight_decompose": True}),
    "locon_dora_in/linear": dict(algo="locon", layer="linear", kw={"weight_decompose": True, "wd_on_out": False}),
    "locon_dora/conv3": dict(algo="locon", layer="conv3", kw={"weight_decompose": True}),
    "locon_tucker/conv3": dict(algo="locon", layer="conv3", kw={}, use_tucker=True),
    "locon_mult/linear": dict(algo="locon", layer="linear", kw={}, multiplier=0.5),
    "lokr_dora/linear": dict(algo="lokr_full", layer="linear", kw={"weight_decompose": True}),
    "lokr_scalar/conv3": dict(algo="lokr_lowrank", layer="conv3", kw={"use_scalar": True}),
    "lokr_mult/linear": dict(algo="lokr_full", layer="linear", kw={}, multiplier=0.5),
    "loha_dora/linear": dict(algo="loha", layer="linear", kw={"weight_decompose": True}),
    "loha_scalar/linear": dict(algo="loha", layer="linear", kw={"use_scalar": True}),
}
REGIMES = ("fp32", "bf16", "autocast_bf16")


def make_base(spec, seed):
    torch.manual_seed(seed)
    if spec["kind"] == "linear":
        return nn.Linear(spec["in_dim"], spec["out_dim"])
    return nn.Conv2d(spec["in_dim"], spec["out_dim"], spec["k"], spec["stride"], spec["pad"])


def perturb(module, seed):
    g = torch.Generator().manual_seed(seed)
    with torch.no_grad():
        for p in module.parameters():
            if float(p.abs().sum()) == 0.0:
                p.copy_(torch.randn(p.shape, generator=g) * 0.05)


def oracle_inputs(algo_key, module):
    """Map reference-module parameters onto the oracle's functional arguments."""
    sd = {k: v for k, v in module.named_parameters()}
    cfg = {"multiplier": 1.0}
    if algo_key == "locon":
        p = {"lora_up.weight": sd["lora_up.weight"], "lora_down.weight": sd["lora_down.weight"]}
        cfg["scale"] = module.scale
        return "locon", p, cfg
    if algo_key.startswith("loha"):
        p = {k: sd[k] for k in ("hada_w1_a", "hada_w1_b", "hada_w2_a", "hada_w2_b")}
        cfg["scale"] = module.scale
        return "loha", p, cfg
    if algo_key.startswith("lokr"):
        p = {k: v for k, v in sd.items() if k.startswith("lokr_")}
        cfg["scale"] = module.scale
        return "lokr", p, cfg
    if algo_key.startswith("ia3"):
        cfg["train_on_input"] = module.train_input
        return "ia3", {"weight": sd["weight"]}, cfg
    if algo_key == "dylora":
        p = {"up_list": list(module.up_list), "down_list": list(module.down_list)}
        cfg["alpha"] = module.alpha
        return "dylora", p, cfg
    raise KeyError(algo_key)


def run_case(algo_key, layer_key, regime, seed, extra_kw=None, multiplier=1.0, use_tucker=False):
    a, spec = ALGOS[algo_key], LAYERS[layer_key]
    a = dict(a, kw={**a["kw"], **(extra_kw or {})})
    base = make_base(spec, seed)
    cls = getattr(lycoris.modules, a["cls"], None) or getattr(
        __import__("lycoris.modules." + {"IA3Module": "ia3", "DyLoraModule": "dylora"}[a["cls"]], fromlist=["x"]), a["cls"])
    torch.manual_seed(seed + 1)
    mod = cls("case", base, multiplier, a["dim"], a["alpha"], 0.0, 0.0, 0.0, use_tucker, **a["kw"])
    perturb(mod, seed + 2)
    if isinstance(getattr(mod, "scalar", None), nn.Parameter):
        with torch.no_grad():
            mod.scalar.fill_(0.7)
    g = torch.Generator().manual_seed(seed + 3)
    x = torch.randn(spec["x"], generator=g)
    ac = None
    if regime == "bf16":
        base.to(torch.bfloat16)
        mod.to(torch.bfloat16)
        x = x.to(torch.bfloat16)
    elif regime == "autocast_bf16":
        base.to(torch.bfloat16)
        x = x.to(torch.bfloat16)
        ac = torch.bfloat16
    mod.apply_to()
    mod.train()
    for p_ in base.parameters():
        p_.requires_grad_(False)

    xr = x.clone().requires_grad_(True)
    random.seed(seed + 4)
    if ac is not None:
        with torch.autocast("cpu", dtype=ac):
            y = base(xr)
    else:
        y = base(xr)
    dy = torch.randn(y.shape, generator=g).to(y.dtype)
    y.backward(dy)
    ref_grads = {k: v.grad.clone() for k, v in mod.named_parameters() if v.grad is not None}
    mod.restore()

    # oracle on the same tensors
    algo, p, cfg = oracle_inputs(algo_key, mod)
    cfg["multiplier"] = multiplier
    cfg["wd_on_out"] = getattr(mod, "wd_on_out", True)
    for extra in ("scalar", "dora_scale", "lora_mid.weight", "hada_t1", "hada_t2"):
        named = dict(mod.named_parameters())
        if extra in named:
            p[extra] = named[extra]
    conv = None
    if spec["kind"] == "conv":
        conv = dict(stride=base.stride, padding=base.padding, dilation=base.dilation, groups=base.groups)
    if algo == "dylora":
        random.seed(seed + 4)
        cfg["b"] = O.draw_dylora_block(mod.block_count)
    yo, dxo, go = O.layer_forward_backward(algo, x, base.weight.detach(), base.bias.detach(), p, cfg, dy, conv, ac)
    assert torch.equal(yo, y.detach()), (algo_key, layer_key, regime, "y", (yo.float() - y.float()).abs().max())
    assert torch.equal(dxo, xr.grad), (algo_key, layer_key, regime, "dx")
    for k, gr in ref_grads.items():
        if algo == "dylora":
            kind, idx = k.split(".")
            og = go[kind][int(idx)]
        else:
            og = go[k]
        assert og is not None and torch.equal(og, gr), (algo_key, layer_key, regime, k)

    case = {
        "weight": base.weight.detach().clone(),
        "bias": base.bias.detach().clone(),
        "x": x.clone(),
        "dy": dy.clone(),
        "y": y.detach().clone(),
        "dx": xr.grad.clone(),
        "params": {k: v.detach().clone() for k, v in mod.named_parameters()},
        "grads": ref_grads,
        "meta": {
            "algo_key": algo_key, "layer": layer_key, "regime": regime, "cls": a["cls"], "dim": a["dim"],
            "alpha": a["alpha"], "kw": a["kw"], "layer_spec": spec, "scale": float(getattr(mod, "scale", 1.0)),
            "multiplier": multiplier, "use_tucker": use_tucker, "wd_on_out": getattr(mod, "wd_on_out", True),
            "dylora_seed": seed + 4, "dylora_b": cfg.get("b"),
        },
    }
    return case


def structure_fixture():
    """Adapter names / classes / state-dict shapes the reference creates on the toy UNet."""
    out = {}

    def sig(net):
        return [[l.lora_name, type(l).__name__, [[k, list(v.shape)] for k, v in l.state_dict().items()]]
                for l in net.loras]

    combos = [
        ("locon", {}), ("loha", {}), ("lokr", {"factor": 8}), ("lokr", {"factor": 4, "decompose_both": True}),
        ("lokr", {"factor": 8, "network_dim_override": 100000}), ("dylora", {"block_size": 2}),
    ]
    for algo, kw in combos:
        kw = dict(kw)
        dim = kw.pop("network_dim_override", 8)
        for preset in ("full", "attn-mlp", "unet-convblock-only", "full-lin"):
            torch.manual_seed(0)
            unet = ToyUNet()
            net = lycoris.kohya.create_network(1.0, dim, 4, None, None, unet, algo=algo, preset=preset,
                                               conv_dim=4, conv_alpha=1, **kw)
            out[f"kohya/{algo}/{json.dumps(kw, sort_keys=True)}/{dim}/{preset}"] = sig(net)
        torch.manual_seed(0)
        unet = ToyUNet()
        lycoris.wrapper.LycorisNetwork.apply_preset(
            {"target_module": ["Linear", "Conv1d", "Conv2d", "Conv3d", "GroupNorm", "LayerNorm"], "target_name": [],
             "module_algo_map": {}, "name_algo_map": {}, "exclude_name": [], "use_fnmatch": False,
             "lora_prefix": "lycoris", "enable_conv": True})
        net = lycoris.create_lycoris(unet, 1.0, dim, 4, algo=algo, conv_dim=4, **kw)
        out[f"wrapper/{algo}/{json.dumps(kw, sort_keys=True)}/{dim}"] = sig(net)
    # per-class / per-name overrides, regex + exclude (docs/Preset.md)
    preset = {
        "enable_conv": True,
        "target_module": ["Transformer2DModel", "ResnetBlock2D"],
        "target_name": ["conv_in", "time_embedding.*"],
        "module_algo_map": {"FeedForward": {"algo": "lokr", "factor": 4, "dim": 100000},
                            "Attention": {"algo": "loha", "dim": 4}},
        "name_algo_map": {"conv_in": {"algo": "locon", "dim": 2}},
        "exclude_name": [".*conv_shortcut"],
        "use_fnmatch": False,
        "lora_prefix": "lycoris",
    }
    lycoris.wrapper.LycorisNetwork.apply_preset(preset)
    torch.manual_seed(0)
    net = lycoris.wrapper.LycorisNetwork(ToyUNet(), 1.0, 8, 4, 1, 1, network_module="locon")
    out["wrapper/algo_map"] = {"preset": preset, "sig": sig(net)}
    return out


def main():
    os.makedirs(OUT, exist_ok=True)
    # 1. integer parity table
    dims = sorted(set(list(range(1, 130)) + [250, 320, 360, 512, 640, 768, 960, 1024, 1280, 1920, 2048, 2560,
                                              2816, 4096, 5120, 10240, 127, 1021]))
    factors = [-1, 1, 2, 4, 6, 8, 12, 16, 64, 100000]
    table = {f"{d},{f}": list(ref_factorization(d, f)) for d in dims for f in factors}
    for key, val in table.items():
        d, f = map(int, key.split(","))
        assert tuple(val) == O.factorization(d, f)
    with open(os.path.join(OUT, "factorization.json"), "w") as fh:
        json.dump(table, fh)

    # 2. numeric layer cases
    n = 0
    for regime in REGIMES:
        cases = {}
        seed = 100
        for algo_key in ALGOS:
            for layer_key in ROUND1_LAYERS:
                seed += 10
                cases[f"{algo_key}/{layer_key}"] = run_case(algo_key, layer_key, regime, seed)
                n += 1
        torch.save(cases, os.path.join(OUT, f"layers_{regime}.pt"))
        opts = {}
        for name, oc in OPTION_CASES.items():
            seed += 10
            opts[name] = run_case(oc["algo"], oc["layer"], regime, seed, oc["kw"], oc.get("multiplier", 1.0),
                                  oc.get("use_tucker", False))
            n += 1
        torch.save(opts, os.path.join(OUT, f"options_{regime}.pt"))
    print(f"{n} layer cases: oracle == reference (bit-exact), fixtures written")

    # 2b. BASELINE.json configs[0]: LoCon dim 4 alpha 1 on nn.Linear(768, 768) through the generic wrapper,
    #     fp32, X ~ N(0,1) [8,77,768], loss = y.float().pow(2).mean()   
    torch.manual_seed(0)
    net_base = nn.Sequential(nn.Linear(768, 768))
    lycoris.wrapper.LycorisNetwork.apply_preset(
        {"target_module": ["Linear"], "target_name": [], "module_algo_map": {}, "name_algo_map": {},
         "exclude_name": [], "use_fnmatch": False, "lora_prefix": "lycoris", "enable_conv": True})
    torch.manual_seed(1)
    net = lycoris.create_lycoris(net_base, 1.0, linear_dim=4, linear_alpha=1, algo="locon")
    net.apply_to()
    perturb(net, 11)
    for p_ in net_base.parameters():
        p_.requires_grad_(False)
    g = torch.Generator().manual_seed(2)
    x = torch.randn(8, 77, 768, generator=g)
    y = net_base(x)
    loss = y.float().pow(2).mean()
    loss.backward()
    lora = net.loras[0]
    # inputs are re-generated from their seeds by the tests (checksums stored); only sample 0 of y is kept
    cfg1 = {
        "weight_sum": float(net_base[0].weight.double().sum()), "x_sum": float(x.double().sum()),
        "bias": net_base[0].bias.detach().clone(),
        "y0": y[0].detach().clone(), "loss": loss.detach().clone(), "lora_name": lora.lora_name,
        "params": {k: v.detach().clone() for k, v in lora.named_parameters()},
        "grads": {k: v.grad.clone() for k, v in lora.named_parameters()},
        "scale": float(lora.scale),
    }
    net.restore()
    yo = O.layer_forward("locon", x, net_base[0].weight.detach(), cfg1["bias"],
                         {k: v for k, v in cfg1["params"].items()}, {"scale": cfg1["scale"], "multiplier": 1.0})
    assert torch.equal(yo, y.detach())
    torch.save(cfg1, os.path.join(OUT, "cfg1_locon_linear768.pt"))
    print("cfg1 fixture written:", cfg1["lora_name"], float(cfg1["loss"]))

    # 3. structure
    with open(os.path.join(OUT, "structure.json"), "w") as fh:
        json.dump(structure_fixture(), fh)
    print("structure fixture written")


if __name__ == "__main__":
    main()
