"""Generate tests/golden/network_side.pt from the REAL reference (build container only; needs
/root/reference): network-level checkpoint round trip, merge and max-norm of SURVEY.md §8f rows 1-2 on the toy
UNet, through `lycoris.kohya` exactly as kohya sd-scripts drives it:

    create_network(...)                              -> state_dict() (the saved checkpoint)
    create_network_from_weights(weights_sd=ckpt)     -> module list (names / classes / shapes) and its own state dict
    network.merge_to(None, unet, ckpt, fp32, "cpu")  -> checksum of every base weight afterwards
    network.apply_max_norm_regularization(v, "cpu")  -> (keys_scaled, mean_norm, max_norm) and the checkpoint afterwards

The tests rebuild the same objects with lycoris_b200.kohya and compare.  Run: ``python oracle/gen_golden_network.py``.
"""

import logging
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(1, "/root/reference")

import lycoris  # noqa: E402  (the reference)
import lycoris.kohya  # noqa: E402

from oracle.toy_models import ToyTextEncoder, ToyUNet  # noqa: E402

logging.getLogger("LyCORIS").setLevel(logging.ERROR)
OUT = os.path.join(ROOT, "tests", "golden", "network_side.pt")

CASES = {
    "lokr_full": dict(dim=100000, alpha=1, kw=dict(algo="lokr", factor=8, preset="full", conv_dim=100000, conv_alpha=1)),
    "lokr_lowrank": dict(dim=2, alpha=1, kw=dict(algo="lokr", factor=4, preset="full", conv_dim=2, conv_alpha=1)),
    "locon": dict(dim=8, alpha=4, kw=dict(algo="locon", preset="full", conv_dim=4, conv_alpha=1)),
    "loha": dict(dim=8, alpha=4, kw=dict(algo="loha", preset="attn-mlp")),
    "locon_dora": dict(dim=4, alpha=2, kw=dict(algo="locon", preset="attn-mlp", dora_wd=True)),
}


def make_network(kohya, case, seed=0):
    torch.manual_seed(seed)
    unet = ToyUNet()
    torch.manual_seed(seed + 1)
    net = kohya.create_network(1.0, case["dim"], case["alpha"], None, None, unet, **case["kw"])
    net.apply_to(None, unet, False, True)  # registers the adapters on the network (kohya calls this next)
    g = torch.Generator().manual_seed(seed + 2)
    with torch.no_grad():
        for p in net.parameters():
            if float(p.abs().sum()) == 0.0:
                p.copy_(torch.randn(p.shape, generator=g) * 0.05)
    return unet, net


def base_checksums(unet):
    return checksums(dict(unet.named_parameters()))


def checksums(sd):
    """(sum, sum of magnitudes) per tensor in float64 — enough to pin a derived state dict without storing it"""
    return {k: [float(v.detach().double().sum()), float(v.detach().double().abs().sum())] for k, v in sd.items()}


def sig(loras):
    return [[l.lora_name, type(l).__name__, [[k, list(v.shape)] for k, v in l.state_dict().items()]] for l in loras]


def snap(sd):
    return {k: v.detach().clone() for k, v in sd.items()}


def run(case):
    out = {"case": case}
    unet, net = make_network(lycoris.kohya, case)
    assert len(list(net.parameters())) > 0
    ckpt = snap(net.state_dict())
    out["checkpoint"] = ckpt
    out["modules"] = sig(net.loras)

    # from-weights on a fresh copy of the base model
    torch.manual_seed(0)
    unet2 = ToyUNet()
    net2, sd2 = lycoris.kohya.create_network_from_weights(1.0, None, None, None, unet2, weights_sd=snap(ckpt))
    # NB `net2.loras` still holds what the constructor built under the class-level preset of the moment; the loaded
    # adapters are `unet_loras` (they become `.loras` in apply_to)
    out["rebuilt_modules"] = sig(net2.unet_loras)
    # kohya's sequence: apply_to registers the rebuilt adapters, then load_state_dict(weights_sd) runs
    net2.apply_to(None, unet2, False, True)
    info = net2.load_state_dict(snap(ckpt), False)
    out["rebuilt_missing"] = sorted(info.missing_keys)
    out["rebuilt_unexpected"] = sorted(info.unexpected_keys)
    out["rebuilt_checkpoint"] = checksums(net2.state_dict())

    # merge into the base weights
    torch.manual_seed(0)
    unet3 = ToyUNet()
    net3, _ = lycoris.kohya.create_network_from_weights(1.0, None, None, None, unet3, weights_sd=snap(ckpt),
                                                         for_inference=True)
    before = base_checksums(unet3)
    net3.merge_to(None, unet3, snap(ckpt), torch.float32, "cpu")
    after = base_checksums(unet3)
    out["merge_changed"] = sorted(k for k in after if after[k] != before[k])
    out["merge_checksums"] = {k: after[k] for k in out["merge_changed"]}

    # max-norm regularisation on the trained network
    unet4, net4 = make_network(lycoris.kohya, case)
    norms = []
    with torch.no_grad():
        for l in net4.loras:
            if hasattr(l, "get_diff_weight"):
                try:
                    norms.append(float(l.get_diff_weight(1.0)[0].norm()))
                except Exception:  # noqa: BLE001
                    pass
    limit = sorted(norms)[len(norms) // 2]  # the median norm: about half of the modules get clamped
    keys_scaled, mean_norm, max_norm = net4.apply_max_norm_regularization(limit, "cpu")
    out["max_norm"] = {"limit": limit, "keys_scaled": int(keys_scaled), "mean_norm": float(mean_norm),
                       "max_norm": float(max_norm), "checkpoint": checksums(net4.state_dict())}
    return out


def run_text_encoders():
    """Adapters on text encoders: one encoder -> `lora_te_*`, a list -> `lora_te1_*`,
    `lora_te2_*`; trained checkpoint, from-weights rebuild, and the unet-only / te-only apply_to switches."""
    out = {}
    for tag, n_te in (("single", 1), ("pair", 2)):
        torch.manual_seed(0)
        unet = ToyUNet()
        tes = [ToyTextEncoder(dim=32 + 16 * i) for i in range(n_te)]
        te_arg = tes[0] if n_te == 1 else tes
        torch.manual_seed(1)
        net = lycoris.kohya.create_network(1.0, 4, 2, None, te_arg, unet, algo="lokr", factor=4, preset="attn-mlp")
        rec = {"te_modules": sig(net.text_encoder_loras), "unet_modules": len(net.unet_loras)}
        net.apply_to(te_arg, unet, True, True)
        g = torch.Generator().manual_seed(2)
        with torch.no_grad():
            for p in net.parameters():
                if float(p.abs().sum()) == 0.0:
                    p.copy_(torch.randn(p.shape, generator=g) * 0.05)
        ckpt = snap(net.state_dict())
        rec["checkpoint"] = ckpt
        # forward of the wrapped text encoder(s) with the trained adapters (fp32, CPU): parity target for the GPU tests
        ids = torch.arange(12).view(2, 6) % 64
        with torch.no_grad():
            rec["ids"] = ids
            rec["te_out"] = [te(ids).clone() for te in tes]
        net.restore()
        torch.manual_seed(0)
        unet2 = ToyUNet()
        tes2 = [ToyTextEncoder(dim=32 + 16 * i) for i in range(n_te)]
        te2 = tes2[0] if n_te == 1 else tes2
        net2, _ = lycoris.kohya.create_network_from_weights(1.0, None, None, te2, unet2, weights_sd=snap(ckpt))
        rec["rebuilt_te_modules"] = sig(net2.text_encoder_loras)
        rec["rebuilt_unet_modules"] = len(net2.unet_loras)
        net2.apply_to(te2, unet2, True, False)  # text encoder only
        rec["te_only_keys"] = list(net2.state_dict().keys())
        out[tag] = rec
    return out


def run_optimizer_groups():
    """prepare_optimizer_params as kohya calls it: group count, lr, parameter count and numel per group, and the
    lr descriptions, for plain / LoRA+ / text-encoder-specific learning rates."""
    out = {}
    torch.manual_seed(0)
    unet = ToyUNet()
    te = ToyTextEncoder()
    torch.manual_seed(1)
    net = lycoris.kohya.create_network(1.0, 4, 2, None, te, unet, algo="locon", preset="attn-mlp")
    net.apply_to(te, unet, True, True)

    def groups(res):
        params, descriptions = res if isinstance(res, tuple) else (res, None)
        return {"groups": [[float(g["lr"]), len(list(g["params"])), int(sum(p.numel() for p in g["params"]))]
                           for g in params], "descriptions": descriptions}

    out["plain"] = groups(net.prepare_optimizer_params(5e-5, 1e-4, 2e-4))
    out["unet_only_lr"] = groups(net.prepare_optimizer_params(None, 1e-4, None))
    net.set_loraplus_lr_ratio(4.0, None, None)
    out["loraplus"] = groups(net.prepare_optimizer_params(5e-5, 1e-4, 2e-4))
    net.set_loraplus_lr_ratio(None, 8.0, 2.0)
    out["loraplus_split"] = groups(net.prepare_optimizer_params(5e-5, 1e-4, 2e-4))
    net.restore()
    return out


def run_fnmatch_preset():
    """name / module maps with shell-style patterns (use_fnmatch) through the generic wrapper"""
    preset = {
        "enable_conv": True, "target_module": ["Transformer2DModel"], "target_name": ["conv_*", "*time_emb_proj"],
        "module_algo_map": {"FeedForward": {"algo": "lokr", "factor": 4, "dim": 100000}},
        "name_algo_map": {"*attn2.to_?": {"algo": "loha", "dim": 2}, "conv_in": {"algo": "locon", "dim": 2}},
        "exclude_name": ["*to_out*"], "use_fnmatch": True, "lora_prefix": "lycoris",
    }
    lycoris.wrapper.LycorisNetwork.apply_preset(dict(preset))
    torch.manual_seed(0)
    net = lycoris.wrapper.LycorisNetwork(ToyUNet(), 1.0, 8, 4, 1, 1, network_module="locon")
    lycoris.wrapper.LycorisNetwork.apply_preset(dict(GENERIC_PRESET))
    return {"preset": preset, "sig": sig(net.loras)}


GENERIC_PRESET = {
    "enable_conv": True, "target_module": ["Linear", "Conv2d"], "target_name": [], "module_algo_map": {},
    "name_algo_map": {}, "exclude_name": [], "use_fnmatch": False, "lora_prefix": "lycoris",
}


def run_generic_wrapper():
    """The model-agnostic API (`lycoris.wrapper`): create_lycoris -> checkpoint -> create_lycoris_from_weights,
    on-the-fly merge / restore (inference-time merge without keeping the adapters attached), and the
    torch parametrization entry point `Module.parametrize`."""
    out = {}
    for algo, kw in (("locon", dict(conv_dim=4, conv_alpha=1)), ("lokr", dict(factor=4))):
        lycoris.wrapper.LycorisNetwork.apply_preset(dict(GENERIC_PRESET))
        torch.manual_seed(0)
        unet = ToyUNet()
        torch.manual_seed(1)
        net = lycoris.create_lycoris(unet, 1.0, linear_dim=4, linear_alpha=2, algo=algo, **kw)
        net.apply_to()
        g = torch.Generator().manual_seed(2)
        with torch.no_grad():
            for p in net.parameters():
                if float(p.abs().sum()) == 0.0:
                    p.copy_(torch.randn(p.shape, generator=g) * 0.05)
        ckpt = snap(net.state_dict())
        rec = {"kw": kw, "modules": sig(net.loras), "checkpoint": ckpt}
        net.restore()
        torch.manual_seed(0)
        unet2 = ToyUNet()
        net2, _ = lycoris.wrapper.create_lycoris_from_weights(0.8, None, unet2, weights_sd=snap(ckpt))
        rec["rebuilt_modules"] = sig(net2.loras)
        rec["rebuilt_algo_table"] = dict(net2.algo_table)
        rec["rebuilt_multipliers"] = sorted({float(l.multiplier) for l in net2.loras})
        before = base_checksums(unet2)
        net2.onfly_merge(0.8)
        merged = base_checksums(unet2)
        rec["onfly_changed"] = sorted(k for k in merged if merged[k] != before[k])
        rec["onfly_checksums"] = {k: merged[k] for k in rec["onfly_changed"]}
        net2.onfly_restore()
        restored = base_checksums(unet2)
        # NB on the CPU `cached_org_weight = org_weight.data.cpu()` ALIASES the live weight (Tensor.cpu() is a no-op
        # there), so the reference's onfly_restore gives back the merged values; on a GPU it restores.  Recorded as is.
        rec["onfly_restored_exactly"] = restored == before
        rec["onfly_after_restore"] = {k: restored[k] for k in rec["onfly_changed"]}
        out[algo] = rec

    # parametrization of a bare weight tensor (LycorisBaseModule.parametrize, base.py:199-247)
    par = {}
    for name, cls, args, kw in (
        ("locon", lycoris.modules.LoConModule, (1.0, 4, 2.0), {}),
        ("lokr", lycoris.modules.LokrModule, (1.0, 2, 1.0), {"factor": 4}),
        ("loha_conv", lycoris.modules.LohaModule, (1.0, 4, 2.0), {}),
    ):
        torch.manual_seed(3)
        host = torch.nn.Conv2d(8, 16, 3) if name.endswith("conv") else torch.nn.Linear(24, 40)
        w0 = host.weight.detach().clone()
        torch.manual_seed(4)
        mod = cls.parametrize(host, "weight", *args, **kw)
        g = torch.Generator().manual_seed(5)
        with torch.no_grad():
            for p in mod.parameters():
                if float(p.abs().sum()) == 0.0:
                    p.copy_(torch.randn(p.shape, generator=g) * 0.05)
        par[name] = {"w0": w0, "params": snap(dict(mod.named_parameters())),
                     "weight": host.weight.detach().clone(),
                     "param_names": sorted(n for n, _ in host.named_parameters())}
    out["parametrize"] = par
    return out


def main():
    cases = {name: run(c) for name, c in CASES.items()}
    cases["text_encoders"] = run_text_encoders()
    cases["generic_wrapper"] = run_generic_wrapper()
    cases["optimizer_groups"] = run_optimizer_groups()
    cases["fnmatch_preset"] = run_fnmatch_preset()
    torch.save(cases, OUT)
    # the public preset tables (lycoris/config.py): data the drop-in has to reproduce key for key
    import json

    from lycoris.config import PRESET

    with open(os.path.join(ROOT, "tests", "golden", "presets.json"), "w") as fh:
        json.dump(PRESET, fh, indent=1, sort_keys=True)
    print(f"{len(cases)} network cases -> {OUT} ({os.path.getsize(OUT) / 1024:.0f} KiB)")
    for k, v in cases.items():
        if k == "text_encoders":
            print("  ", k, {t: len(r["te_modules"]) for t, r in v.items()})
            continue
        if k in ("optimizer_groups", "fnmatch_preset"):
            print("  ", k, {t: r for t, r in v.items() if t != "sig"} if k == "optimizer_groups" else len(v["sig"]))
            continue
        if k == "generic_wrapper":
            print("  ", k, {t: (len(r["modules"]), len(r["onfly_changed"]), r["onfly_restored_exactly"])
                            for t, r in v.items() if t != "parametrize"}, sorted(v["parametrize"]))
            continue
        print("  ", k, len(v["modules"]), "modules; merged", len(v["merge_changed"]), "base tensors; max-norm scaled",
              v["max_norm"]["keys_scaled"], "missing", len(v["rebuilt_missing"]), "unexpected", len(v["rebuilt_unexpected"]))


if __name__ == "__main__":
    main()
