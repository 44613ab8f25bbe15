"""Generate tests/golden/weights_side.pt from the REAL reference (build container only; needs
/root/reference): the checkpoint wire format and the weight-only entry points of SURVEY.md §8f rows 1-2,
plus bypass-mode forward/backward (row 4), per adapter type on a Linear and a 3x3 Conv2d:

    state_dict()                      key names, alpha buffer, scalar folded into the first factor
    get_diff_weight(0.7)              incl. the LoHa/LoKr double-scale quirk
    get_merged_weight(0.7)
    merge_to(0.5)                     base weight (and bias for IA3) after the in-place merge
    apply_max_norm(limit)             (scaled?, norm) and the state dict afterwards
    bypass_mode=True fwd+bwd          y, dx, parameter grads (fp32)

The product classes are rebuilt FROM THE REFERENCE'S state dict (make_module_from_state_dict) by the
tests, so this also pins the from-weights round trip.  Run: ``python oracle/gen_golden_weights.py``.
"""

import logging
import os
import random
import sys

import torch
import torch.nn as nn

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(1, "/root/reference")

import lycoris  # noqa: E402  (the reference)
import lycoris.modules  # noqa: E402
from lycoris.modules.dylora import DyLoraModule  # noqa: E402
from lycoris.modules.ia3 import IA3Module  # noqa: E402

from oracle.gen_golden import LAYERS, make_base, perturb  # noqa: E402

logging.getLogger("LyCORIS").setLevel(logging.ERROR)
OUT = os.path.join(ROOT, "tests", "golden", "weights_side.pt")

CLS = {
    "LoConModule": lycoris.modules.LoConModule, "LohaModule": lycoris.modules.LohaModule,
    "LokrModule": lycoris.modules.LokrModule, "IA3Module": IA3Module, "DyLoraModule": DyLoraModule,
}
CASES = {
    "locon": dict(cls="LoConModule", dim=4, alpha=2.0, kw={}),
    "locon_scalar": dict(cls="LoConModule", dim=4, alpha=2.0, kw={"use_scalar": True}),
    "locon_dora": dict(cls="LoConModule", dim=4, alpha=2.0, kw={"weight_decompose": True}),
    "locon_tucker": dict(cls="LoConModule", dim=4, alpha=2.0, kw={}, use_tucker=True, layers=("conv3",)),
    "loha": dict(cls="LohaModule", dim=4, alpha=2.0, kw={}),
    "loha_scalar": dict(cls="LohaModule", dim=4, alpha=2.0, kw={"use_scalar": True}),
    "lokr_full": dict(cls="LokrModule", dim=100000, alpha=1.0, kw={"factor": 4}),
    "lokr_lowrank": dict(cls="LokrModule", dim=2, alpha=1.0, kw={"factor": 4}),
    "lokr_both": dict(cls="LokrModule", dim=1, alpha=1.0, kw={"factor": 8, "decompose_both": True}),
    "lokr_dora": dict(cls="LokrModule", dim=2, alpha=1.0, kw={"factor": 4, "weight_decompose": True}),
    "ia3_out": dict(cls="IA3Module", dim=4, alpha=1.0, kw={"train_on_input": False}),
    "ia3_in": dict(cls="IA3Module", dim=4, alpha=1.0, kw={"train_on_input": True}),
    "dylora": dict(cls="DyLoraModule", dim=8, alpha=4.0, kw={"block_size": 2}),
    # rank dropout: Bernoulli mask over dW rows, drawn from torch's global RNG in training mode
    "locon_rankdrop": dict(cls="LoConModule", dim=4, alpha=2.0, kw={}, rank_dropout=0.5),
    "locon_rankdrop_scaled": dict(cls="LoConModule", dim=4, alpha=2.0, kw={"rank_dropout_scale": True}, rank_dropout=0.5),
    "loha_rankdrop": dict(cls="LohaModule", dim=4, alpha=2.0, kw={}, rank_dropout=0.5),
    "lokr_rankdrop": dict(cls="LokrModule", dim=100000, alpha=1.0, kw={"factor": 4}, rank_dropout=0.5),
}


def build(c, layer_key, seed, bypass=None):
    base = make_base(LAYERS[layer_key], seed)
    torch.manual_seed(seed + 1)
    kw = dict(c["kw"])
    if bypass is not None:
        kw["bypass_mode"] = bypass
    mod = CLS[c["cls"]]("case", base, 1.0, c["dim"], c["alpha"], 0.0, c.get("rank_dropout", 0.0), 0.0,
                        c.get("use_tucker", False), **kw)
    perturb(mod, seed + 2)
    if isinstance(getattr(mod, "scalar", None), nn.Parameter):
        with torch.no_grad():
            mod.scalar.fill_(0.7)
    return base, mod


def snap(sd):
    return {k: v.detach().clone() for k, v in sd.items()}


def checksums(sd):
    """(sum, sum of magnitudes) per tensor in float64: pins a derived state dict without storing it again"""
    return {k: [float(v.detach().double().sum()), float(v.detach().double().abs().sum())] for k, v in sd.items()}


def try_call(fn, *a, **k):
    try:
        return fn(*a, **k)
    except Exception as e:  # noqa: BLE001 - the reference's own error type is part of the contract
        return {"raises": type(e).__name__}


def run(name, c, layer_key, seed):
    out = {"meta": dict(cls=c["cls"], dim=c["dim"], alpha=c["alpha"], kw=c["kw"], use_tucker=c.get("use_tucker", False),
                        layer=layer_key, layer_spec=LAYERS[layer_key], seed=seed, rank_dropout=c.get("rank_dropout", 0.0))}
    base, mod = build(c, layer_key, seed)
    out["weight"] = base.weight.detach().clone()
    out["bias"] = base.bias.detach().clone()
    out["params"] = snap(dict(mod.named_parameters()))
    out["state_dict"] = snap(mod.state_dict())
    with torch.no_grad():
        mod.eval()
        random.seed(seed + 5)  # DyLoRA draws its rank from Python's RNG when none is given
        d = try_call(mod.get_diff_weight, 0.7)
        out["diff_0p7"] = d if isinstance(d, dict) else [None if t is None else t.clone() for t in d]
        random.seed(seed + 6)
        m = try_call(mod.get_merged_weight, 0.7)
        out["merged_0p7"] = m if isinstance(m, dict) else [None if t is None else t.clone() for t in m]
        random.seed(seed + 7)
        r = try_call(mod.merge_to, 0.5)
        out["merge_to_0p5"] = r if isinstance(r, dict) else {"weight": base.weight.detach().clone(),
                                                               "bias": base.bias.detach().clone()}
    if c.get("rank_dropout", 0.0):
        # training mode: the mask comes from torch's RNG — same seed, same rows dropped
        mod.train()
        torch.manual_seed(seed + 9)
        with torch.no_grad():
            out["train_diff"] = mod.get_diff_weight(1.0)[0].clone()
        mod.eval()
    # the module rebuilt from its own checkpoint (create_network_from_weights' per-layer step)
    if c["cls"] not in ("IA3Module",):  # quirk 1: IA3's loader has an arity bug in the reference
        from lycoris.modules import get_module, make_module

        sd = {f"case.{k}": v.clone() for k, v in out["state_dict"].items()}
        fresh = make_base(LAYERS[layer_key], seed)

        def rebuild():
            cls_, weights = get_module(sd, "case")
            m2 = make_module(cls_, weights, "case", fresh)
            m2.eval()
            random.seed(seed + 5)
            with torch.no_grad():
                return {"cls": type(m2).__name__, "diff_0p7": m2.get_diff_weight(0.7)[0].clone(),
                        "state_dict": checksums(m2.state_dict())}

        out["rebuilt"] = try_call(rebuild)
    # max-norm on a fresh copy; the limit is chosen below the current norm so the clamp engages
    if hasattr(mod, "apply_max_norm"):
        base, mod = build(c, layer_key, seed)
        mod.eval()
        with torch.no_grad():
            cur = float(mod.get_diff_weight(1.0)[0].norm())
        mod.train()  # kohya calls it on the training network: rank dropout (if any) draws from torch's RNG
        limit = max(cur * 0.5, 1e-4)
        torch.manual_seed(seed + 8)
        res = try_call(mod.apply_max_norm, limit, None)
        if isinstance(res, dict):
            out["max_norm"] = res
        else:
            scaled, norm = res
            out["max_norm"] = {"limit": limit, "scaled": bool(scaled), "norm": None if norm is None else torch.as_tensor(norm).detach().clone(),
                               "state_dict": checksums(mod.state_dict()), "params": snap(dict(mod.named_parameters()))}
    # bypass mode, fp32, training
    base, mod = build(c, layer_key, seed, bypass=True)
    mod.apply_to()
    mod.train()
    for p_ in base.parameters():
        p_.requires_grad_(False)
    g = torch.Generator().manual_seed(seed + 3)
    x = torch.randn(LAYERS[layer_key]["x"], generator=g)
    xr = x.clone().requires_grad_(True)
    random.seed(seed + 4)
    torch.manual_seed(seed + 10)  # rank dropout in bypass mode masks the rank dimension from torch's RNG
    y = try_call(base, xr)
    if isinstance(y, dict):
        out["bypass"] = y
    else:
        dy = torch.randn(y.shape, generator=g)
        y.backward(dy)
        out["bypass"] = {"x": x, "dy": dy, "y": y.detach().clone(), "dx": xr.grad.clone(), "rand_seed": seed + 4,
                         "grads": {k: v.grad.clone() for k, v in mod.named_parameters() if v.grad is not None}}
    mod.restore()
    return out


def main():
    cases = {}
    seed = 5000
    for name, c in CASES.items():
        for layer_key in c.get("layers", ("linear", "conv3")):
            seed += 10
            cases[f"{name}/{layer_key}"] = run(name, c, layer_key, seed)
    torch.save(cases, OUT)
    print(f"{len(cases)} weight-side cases written to {OUT} ({os.path.getsize(OUT) / 1024:.0f} KiB)")
    for k, v in cases.items():
        flags = [t for t in ("diff_0p7", "merged_0p7", "merge_to_0p5", "max_norm", "bypass", "rebuilt")
                 if isinstance(v.get(t), dict) and "raises" in v[t]]
        if flags:
            print("  ", k, "reference raises in", {t: v[t]["raises"] for t in flags})


if __name__ == "__main__":
    main()
