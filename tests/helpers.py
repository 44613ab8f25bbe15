"""Shared helpers for the parity tests: build product modules from golden cases, run the oracle."""
import os
import random

import torch
import torch.nn as nn

from conftest import GOLDEN

_CACHE = {}


def load_cases(regime, kind="layers"):
    key = (kind, regime)
    if key not in _CACHE:
        _CACHE[key] = torch.load(os.path.join(GOLDEN, f"{kind}_{regime}.pt"), weights_only=False)
    return _CACHE[key]


def case_ids(regime, kind="layers"):
    return sorted(load_cases(regime, kind).keys())


def build_base(case, device="cpu"):
    spec = case["meta"]["layer_spec"]
    if spec["kind"] == "linear":
        base = nn.Linear(spec["in_dim"], spec["out_dim"])
    else:
        base = nn.Conv2d(spec["in_dim"], spec["out_dim"], spec["k"], spec["stride"], spec["pad"])
    base.weight.data = case["weight"].clone()
    base.bias.data = case["bias"].clone()
    for p in base.parameters():
        p.requires_grad_(False)
    return base.to(device)


def build_product_module(case, base):
    """lycoris_b200 module with the golden case's parameters (same constructor call as the reference)."""
    import lycoris_b200.modules as M
    from lycoris_b200.modules import dylora, ia3  # noqa: F401

    meta = case["meta"]
    cls = {"LoConModule": M.LoConModule, "LohaModule": M.LohaModule, "LokrModule": M.LokrModule,
           "IA3Module": M.IA3Module, "DyLoraModule": M.DyLoraModule}[meta["cls"]]
    mod = cls("case", base, meta.get("multiplier", 1.0), meta["dim"], meta["alpha"], 0.0, meta.get("rank_dropout", 0.0),
              0.0, meta.get("use_tucker", False), **meta["kw"])
    own = dict(mod.named_parameters())
    assert set(own) == set(case["params"]), (sorted(own), sorted(case["params"]))
    with torch.no_grad():
        for k, v in case["params"].items():
            assert own[k].shape == v.shape, (k, own[k].shape, v.shape)
            own[k].data = v.clone().to(own[k].device)
    return mod


def oracle_args(case, device="cpu"):
    """(algo, params, cfg, conv) for oracle.lyco_oracle.layer_forward_backward."""
    meta = case["meta"]
    key = meta["algo_key"]
    p = {k: v.to(device) for k, v in case["params"].items()}
    cfg = {"multiplier": meta.get("multiplier", 1.0), "scale": meta["scale"], "wd_on_out": meta.get("wd_on_out", True)}
    if key == "locon":
        algo = "locon"
    elif key.startswith("loha"):
        algo = "loha"
    elif key.startswith("lokr"):
        algo = "lokr"
    elif key.startswith("ia3"):
        algo = "ia3"
        cfg["train_on_input"] = meta["kw"]["train_on_input"]
    else:
        algo = "dylora"
        nb = meta["dim"] // meta["kw"]["block_size"]
        p = {"up_list": [p[f"up_list.{i}"] for i in range(nb)], "down_list": [p[f"down_list.{i}"] for i in range(nb)]}
        cfg["alpha"] = torch.tensor(meta["alpha"], device=device)
        cfg["b"] = meta["dylora_b"]
    spec = meta["layer_spec"]
    conv = None
    if spec["kind"] == "conv":
        conv = dict(stride=(spec["stride"],) * 2, padding=(spec["pad"],) * 2, dilation=(1, 1), groups=1)
    return algo, p, cfg, conv


def flat_grads(case_grads):
    return {k: v for k, v in case_grads.items()}


def rel_err(a, b):
    a, b = a.float(), b.float()
    return float((a - b).norm() / (b.norm() + 1e-12))


def seed_dylora(case):
    random.seed(case["meta"]["dylora_seed"])


def oracle_patch_network(net):
    """Route every wrapped layer of a (not applied) lycoris_b200 network through the ORACLE's
    per-layer forward using the network's own parameter tensors — a reference-equivalent eager
    network on whatever device the model lives on.  Returns an undo callable."""
    import random as _random

    from oracle import lyco_oracle as O

    undo = []
    for lora in net.loras:
        org = lora.org_module[0]
        conv = None
        if lora.module_type.startswith("conv"):
            conv = dict(stride=org.stride, padding=org.padding, dilation=org.dilation, groups=org.groups)
        cls = type(lora).__name__

        def fwd(x, _l=lora, _o=org, _c=conv, _cls=cls):
            p = dict(_l.named_parameters())
            cfg = {"multiplier": _l.multiplier, "scale": getattr(_l, "scale", 1.0)}
            if _cls == "LoConModule":
                algo = "locon"
            elif _cls == "LohaModule":
                algo = "loha"
            elif _cls == "LokrModule":
                algo = "lokr"
            elif _cls == "IA3Module":
                algo = "ia3"
                cfg["train_on_input"] = _l.train_input
            else:
                algo = "dylora"
                p = {"up_list": list(_l.up_list), "down_list": list(_l.down_list)}
                cfg["alpha"] = _l.alpha
                cfg["b"] = _random.randint(0, _l.block_count - 1)
            return O.layer_forward(algo, x, _o.weight, _o.bias, p, cfg, _c)

        saved = org.forward
        org.forward = fwd
        undo.append((org, saved))

    def restore():
        for org, saved in undo:
            org.forward = saved

    return restore


def load_cfg1():
    """BASELINE.json configs[0] fixture; weight / x are re-created from their seeds and checked against
    the stored checksums (the fixture keeps only what cannot be regenerated: reference outputs)."""
    c = torch.load(os.path.join(GOLDEN, "cfg1_locon_linear768.pt"), weights_only=False)
    torch.manual_seed(0)
    lin = nn.Linear(768, 768)
    g = torch.Generator().manual_seed(2)
    x = torch.randn(8, 77, 768, generator=g)
    assert abs(float(lin.weight.double().sum()) - c["weight_sum"]) < 1e-9, "seeded weight differs from the fixture's"
    assert abs(float(x.double().sum()) - c["x_sum"]) < 1e-9, "seeded input differs from the fixture's"
    assert torch.equal(lin.bias.detach(), c["bias"])
    c["weight"], c["x"] = lin.weight.detach().clone(), x
    return c
