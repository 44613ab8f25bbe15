"""Network-level checkpoint round trip, merge and max-norm through ``lycoris_b200.kohya`` against the REAL
reference's outputs on the toy UNet (tests/golden/network_side.pt, written by oracle/gen_golden_network.py):
SURVEY.md §8f rows 1-2 as kohya sd-scripts drives them.  Host-side logic, runs on the CPU."""
import logging
import os

import pytest
import torch

from conftest import GOLDEN

CASES = torch.load(os.path.join(GOLDEN, "network_side.pt"), weights_only=False)
IDS = sorted(k for k in CASES if "checkpoint" in CASES[k] and "max_norm" in CASES[k])
logging.getLogger("LyCORIS").setLevel(logging.ERROR)


def _toy(seed=0):
    from oracle.toy_models import ToyUNet

    torch.manual_seed(seed)
    return ToyUNet()


def _make_network(case, seed=0):
    """Same construction sequence as oracle/gen_golden_network.py:make_network, with the product's kohya module."""
    import lycoris_b200.kohya as kohya

    unet = _toy(seed)
    torch.manual_seed(seed + 1)
    net = kohya.create_network(1.0, case["dim"], case["alpha"], None, None, unet, **case["kw"])
    net.apply_to(None, unet, False, True)
    g = torch.Generator().manual_seed(seed + 2)
    with torch.no_grad():
        for p in net.parameters():
            if float(p.abs().sum()) == 0.0:
                p.copy_(torch.randn(p.shape, generator=g) * 0.05)
    return unet, net


def _sig(loras):
    return [[l.lora_name, type(l).__name__, [[k, list(v.shape)] for k, v in l.state_dict().items()]] for l in loras]


def _checksums(sd):
    return {k: [float(v.detach().double().sum()), float(v.detach().double().abs().sum())] for k, v in sd.items()}


def _close(a, b, what, rel=1e-5):
    assert set(a) == set(b), (what, sorted(set(a) ^ set(b))[:6])
    for k in b:
        for x, y in zip(a[k], b[k]):
            assert abs(x - y) <= rel * max(1.0, abs(y)), (what, k, x, y)


def _snap(sd):
    return {k: v.detach().clone() for k, v in sd.items()}


@pytest.mark.parametrize("name", IDS)
def test_trained_network_saves_the_reference_checkpoint(name):
    """Same seeds, same construction: every key, shape and value of state_dict() equals the reference's."""
    ref = CASES[name]
    _, net = _make_network(ref["case"])
    assert _sig(net.loras) == ref["modules"]
    sd = net.state_dict()
    assert list(sd.keys()) == list(ref["checkpoint"].keys())
    for k, v in ref["checkpoint"].items():
        assert sd[k].shape == v.shape and sd[k].dtype == v.dtype, k
        assert torch.allclose(sd[k].detach(), v, rtol=1e-6, atol=1e-7), (k, float((sd[k].detach() - v).abs().max()))


@pytest.mark.parametrize("name", IDS)
def test_create_network_from_reference_checkpoint(name):
    """create_network_from_weights on the reference's checkpoint: same module list as the reference's own loader,
    nothing missing / unexpected on load, and the re-exported checkpoint matches."""
    import lycoris_b200.kohya as kohya

    ref = CASES[name]
    unet = _toy()
    net, sd = kohya.create_network_from_weights(1.0, None, None, None, unet, weights_sd=_snap(ref["checkpoint"]))
    assert _sig(net.unet_loras) == ref["rebuilt_modules"]  # (.loras is only rebuilt from them in apply_to)
    net.apply_to(None, unet, False, True)
    assert _sig(net.loras) == ref["rebuilt_modules"]
    info = net.load_state_dict(_snap(ref["checkpoint"]), False)
    assert sorted(info.missing_keys) == ref["rebuilt_missing"]
    assert sorted(info.unexpected_keys) == ref["rebuilt_unexpected"]
    _close(_checksums(net.state_dict()), ref["rebuilt_checkpoint"], f"{name}: re-exported checkpoint")


@pytest.mark.parametrize("name", IDS)
def test_merge_to_base_model(name):
    """for_inference load + merge_to: exactly the base tensors the reference touches end up with its values."""
    import lycoris_b200.kohya as kohya

    ref = CASES[name]
    unet = _toy()
    net, _ = kohya.create_network_from_weights(1.0, None, None, None, unet, weights_sd=_snap(ref["checkpoint"]),
                                               for_inference=True)
    before = _checksums(dict(unet.named_parameters()))
    net.merge_to(None, unet, _snap(ref["checkpoint"]), torch.float32, "cpu")
    after = _checksums(dict(unet.named_parameters()))
    changed = sorted(k for k in after if after[k] != before[k])
    assert changed == ref["merge_changed"]
    _close({k: after[k] for k in changed}, ref["merge_checksums"], f"{name}: merged base weights")


@pytest.mark.parametrize("name", IDS)
def test_max_norm_regularisation(name):
    """kohya's scale_weight_norms hook: same count of scaled modules, same mean / max norm, same parameters after."""
    ref = CASES[name]
    exp = ref["max_norm"]
    _, net = _make_network(ref["case"])
    keys_scaled, mean_norm, max_norm = net.apply_max_norm_regularization(exp["limit"], "cpu")
    assert int(keys_scaled) == exp["keys_scaled"]
    assert abs(float(mean_norm) - exp["mean_norm"]) <= 1e-5 * max(1.0, abs(exp["mean_norm"]))
    assert abs(float(max_norm) - exp["max_norm"]) <= 1e-5 * max(1.0, abs(exp["max_norm"]))
    _close(_checksums(net.state_dict()), exp["checkpoint"], f"{name}: checkpoint after max-norm")


# ----------------------------------------------------------------------------- text encoders
def _text_encoders(n):
    from oracle.toy_models import ToyTextEncoder

    tes = [ToyTextEncoder(dim=32 + 16 * i) for i in range(n)]
    return tes, (tes[0] if n == 1 else tes)


@pytest.mark.parametrize("tag,n_te", [("single", 1), ("pair", 2)])
def test_text_encoder_adapters_match_reference(tag, n_te):
    """One encoder -> `lora_te_*`, a list -> `lora_te1_*` / `lora_te2_*` (CLIPAttention / CLIPMLP discovery); the
    trained checkpoint equals the reference's, the from-weights loader rebuilds the same modules, and
    apply_to(text encoder only) registers only those."""
    import lycoris_b200.kohya as kohya

    ref = CASES["text_encoders"][tag]
    unet = _toy()
    tes, te_arg = _text_encoders(n_te)
    torch.manual_seed(1)
    net = kohya.create_network(1.0, 4, 2, None, te_arg, unet, algo="lokr", factor=4, preset="attn-mlp")
    assert _sig(net.text_encoder_loras) == ref["te_modules"]
    assert len(net.unet_loras) == ref["unet_modules"]
    prefixes = {l.lora_name.split("_text_model")[0] for l in net.text_encoder_loras}
    assert prefixes == ({"lora_te"} if n_te == 1 else {"lora_te1", "lora_te2"})
    net.apply_to(te_arg, unet, True, True)
    g = torch.Generator().manual_seed(2)
    with torch.no_grad():
        for p in net.parameters():
            if float(p.abs().sum()) == 0.0:
                p.copy_(torch.randn(p.shape, generator=g) * 0.05)
    sd = net.state_dict()
    assert list(sd.keys()) == list(ref["checkpoint"].keys())
    for k, v in ref["checkpoint"].items():
        assert torch.allclose(sd[k].detach(), v, rtol=1e-6, atol=1e-7), k
    net.restore()

    unet2 = _toy()
    _, te2 = _text_encoders(n_te)
    net2, _ = kohya.create_network_from_weights(1.0, None, None, te2, unet2, weights_sd=_snap(ref["checkpoint"]))
    assert _sig(net2.text_encoder_loras) == ref["rebuilt_te_modules"]
    assert len(net2.unet_loras) == ref["rebuilt_unet_modules"]
    net2.apply_to(te2, unet2, True, False)
    assert list(net2.state_dict().keys()) == ref["te_only_keys"]
    net2.restore()


# ------------------------------------------------------------------- model-agnostic wrapper (lycoris.wrapper)
GENERIC_PRESET = {
    "enable_conv": True, "target_module": ["Linear", "Conv2d"], "target_name": [], "module_algo_map": {},
    "name_algo_map": {}, "exclude_name": [], "use_fnmatch": False, "lora_prefix": "lycoris",
}


@pytest.mark.parametrize("algo", ["locon", "lokr"])
def test_generic_wrapper_checkpoint_from_weights_and_onfly_merge(algo):
    import lycoris_b200 as L
    from lycoris_b200.wrapper import LycorisNetwork, create_lycoris_from_weights

    ref = CASES["generic_wrapper"][algo]
    LycorisNetwork.apply_preset(dict(GENERIC_PRESET))
    unet = _toy()
    torch.manual_seed(1)
    net = L.create_lycoris(unet, 1.0, linear_dim=4, linear_alpha=2, algo=algo, **ref["kw"])
    net.apply_to()
    g = torch.Generator().manual_seed(2)
    with torch.no_grad():
        for p in net.parameters():
            if float(p.abs().sum()) == 0.0:
                p.copy_(torch.randn(p.shape, generator=g) * 0.05)
    assert _sig(net.loras) == ref["modules"]
    sd = net.state_dict()
    assert list(sd.keys()) == list(ref["checkpoint"].keys())
    for k, v in ref["checkpoint"].items():
        assert torch.allclose(sd[k].detach(), v, rtol=1e-6, atol=1e-7), k
    net.restore()

    unet2 = _toy()
    net2, _ = create_lycoris_from_weights(0.8, None, unet2, weights_sd=_snap(ref["checkpoint"]))
    assert _sig(net2.loras) == ref["rebuilt_modules"]
    assert dict(net2.algo_table) == ref["rebuilt_algo_table"]
    assert sorted({float(l.multiplier) for l in net2.loras}) == ref["rebuilt_multipliers"]
    before = _checksums(dict(unet2.named_parameters()))
    net2.onfly_merge(0.8)
    merged = _checksums(dict(unet2.named_parameters()))
    changed = sorted(k for k in merged if merged[k] != before[k])
    assert changed == ref["onfly_changed"]
    _close({k: merged[k] for k in changed}, ref["onfly_checksums"], f"{algo}: on-the-fly merged weights")
    net2.onfly_restore()
    restored = _checksums(dict(unet2.named_parameters()))
    # same behaviour as the reference on this device, including its CPU aliasing quirk (see the generator)
    _close({k: restored[k] for k in changed}, ref["onfly_after_restore"], f"{algo}: after onfly_restore")


@pytest.mark.parametrize("name", ["locon", "lokr", "loha_conv"])
def test_parametrize_entry_point(name):
    """`Module.parametrize(host, "weight", ...)`: the host's weight becomes W + dW through torch's parametrization
    machinery; same parameter names and the same parametrized weight as the reference."""
    import lycoris_b200.modules as M

    ref = CASES["generic_wrapper"]["parametrize"][name]
    cls, args, kw = {"locon": (M.LoConModule, (1.0, 4, 2.0), {}), "lokr": (M.LokrModule, (1.0, 2, 1.0), {"factor": 4}),
                     "loha_conv": (M.LohaModule, (1.0, 4, 2.0), {})}[name]
    torch.manual_seed(3)
    host = torch.nn.Conv2d(8, 16, 3) if name.endswith("conv") else torch.nn.Linear(24, 40)
    assert torch.equal(host.weight.detach(), ref["w0"])
    torch.manual_seed(4)
    mod = cls.parametrize(host, "weight", *args, **kw)
    own = dict(mod.named_parameters())
    assert set(own) == set(ref["params"])
    with torch.no_grad():
        for k, v in ref["params"].items():
            own[k].copy_(v)
    assert sorted(n for n, _ in host.named_parameters()) == ref["param_names"]
    assert torch.allclose(host.weight.detach(), ref["weight"], rtol=1e-5, atol=1e-6)
    assert not torch.equal(host.weight.detach(), ref["w0"])


def test_optimizer_param_groups_match_reference():
    """prepare_optimizer_params (plain, unet-only, LoRA+ with one ratio, LoRA+ with separate unet / text-encoder
    ratios): same groups, learning rates, parameter counts and descriptions as the reference."""
    import lycoris_b200.kohya as kohya
    from oracle.toy_models import ToyTextEncoder

    ref = CASES["optimizer_groups"]
    unet = _toy()
    te = ToyTextEncoder()
    torch.manual_seed(1)
    net = kohya.create_network(1.0, 4, 2, None, te, unet, algo="locon", preset="attn-mlp")
    net.apply_to(te, unet, True, True)

    def groups(res):
        params, descriptions = res if isinstance(res, tuple) else (res, None)
        return {"groups": [[float(g["lr"]), len(list(g["params"])), int(sum(p.numel() for p in g["params"]))]
                           for g in params], "descriptions": descriptions}

    assert groups(net.prepare_optimizer_params(5e-5, 1e-4, 2e-4)) == ref["plain"]
    assert groups(net.prepare_optimizer_params(None, 1e-4, None)) == ref["unet_only_lr"]
    net.set_loraplus_lr_ratio(4.0, None, None)
    assert groups(net.prepare_optimizer_params(5e-5, 1e-4, 2e-4)) == ref["loraplus"]
    net.set_loraplus_lr_ratio(None, 8.0, 2.0)
    assert groups(net.prepare_optimizer_params(5e-5, 1e-4, 2e-4)) == ref["loraplus_split"]
    net.restore()


def test_fnmatch_preset_selects_the_reference_modules():
    """use_fnmatch: shell-style patterns in target_name / name_algo_map / exclude_name (docs/Preset.md)."""
    from lycoris_b200.wrapper import LycorisNetwork

    ref = CASES["fnmatch_preset"]
    LycorisNetwork.apply_preset(dict(ref["preset"]))
    try:
        torch.manual_seed(0)
        net = LycorisNetwork(_toy(), 1.0, 8, 4, 1, 1, network_module="locon")
    finally:
        LycorisNetwork.apply_preset(dict(GENERIC_PRESET))
    assert _sig(net.loras) == ref["sig"]
    assert {"LoConModule", "LokrModule"} <= {s[1] for s in ref["sig"]}  # the module_algo_map override took effect
