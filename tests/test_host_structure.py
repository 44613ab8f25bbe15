"""CPU: host-side drop-in behaviour — adapter discovery, names, classes, state-dict keys/shapes
(bit-exact against fixtures produced by the reference wrapper / kohya adapter), forward patching
and stacking linkage, checkpoint round trip.  No compute calls (no GPU here)."""
import json
import os

import pytest
import torch
import torch.nn as nn

import lycoris_b200 as L
from conftest import GOLDEN
from lycoris_b200.kohya import LycorisNetworkKohya, create_network, create_network_from_weights
from lycoris_b200.wrapper import LycorisNetwork, create_lycoris, create_lycoris_from_weights
from oracle.toy_models import ToyUNet

STRUCT = json.load(open(os.path.join(GOLDEN, "structure.json")))


def sig(net):
    return [[l.lora_name, type(l).__name__, [[k, list(v.shape)] for k, v in l.state_dict().items()]] for l in net.loras]


@pytest.mark.parametrize("key", [k for k in STRUCT if k.startswith("kohya/")])
def test_kohya_network_structure_matches_reference(key):
    _, algo, kw, dim, preset = key.split("/")
    kw = json.loads(kw)
    torch.manual_seed(0)
    net = create_network(1.0, int(dim), 4, None, None, ToyUNet(), algo=algo, preset=preset, conv_dim=4, conv_alpha=1, **kw)
    assert sig(net) == STRUCT[key]
    assert all(n.startswith("lora_unet_") for n, *_ in sig(net))


@pytest.mark.parametrize("key", [k for k in STRUCT if k.startswith("wrapper/") and k != "wrapper/algo_map"])
def test_generic_wrapper_structure_matches_reference(key):
    _, algo, kw, dim = key.split("/")
    kw = json.loads(kw)
    torch.manual_seed(0)
    net = create_lycoris(ToyUNet(), 1.0, int(dim), 4, algo=algo, conv_dim=4, **kw)
    assert sig(net) == STRUCT[key]


def test_algo_maps_regex_and_exclude():
    entry = STRUCT["wrapper/algo_map"]
    LycorisNetwork.apply_preset(entry["preset"])
    torch.manual_seed(0)
    net = LycorisNetwork(ToyUNet(), 1.0, 8, 4, 1, 1, network_module="locon")
    assert sig(net) == entry["sig"]
    classes = {n: c for n, c, _ in sig(net)}
    assert any(c == "LokrModule" for c in classes.values()) and any(c == "LohaModule" for c in classes.values())
    # NB exclude_name only filters the top-level walk (wrapper.py:419-423): layers reached through a
    # class-matched parent (ResnetBlock2D.conv_shortcut) are still wrapped — same as the reference.


def test_unknown_preset_key_raises_keyerror():
    with pytest.raises(KeyError):
        LycorisNetwork.apply_preset({"no_such_key": 1})


def test_ia3_is_registered_and_out_of_scope_algos_are_explicit():
    net = create_lycoris(nn.Sequential(nn.Linear(8, 8)), 1.0, 4, 1, algo="ia3")
    assert type(net.loras[0]).__name__ == "IA3Module"
    with pytest.raises(KeyError, match="not part of the B200 adapter engine"):
        create_lycoris(nn.Sequential(nn.Linear(8, 8)), 1.0, 4, 1, algo="boft")


def test_dylora_rank_must_divide():
    with pytest.raises(AssertionError):
        L.DyLoraModule("x", nn.Linear(8, 8), 1.0, 6, 1, block_size=4)


def test_unsupported_module_type_raises():
    with pytest.raises(ValueError):
        L.LoConModule("x", nn.LayerNorm(8), 1.0, 4, 1)


def test_apply_restore_and_stacking_linkage():
    base = nn.Linear(8, 8)
    orig = base.forward
    a = L.LoConModule("a", base, 1.0, 2, 1)
    b = L.LokrModule("b", base, 1.0, 2, 1)
    a.apply_to()
    assert base.forward == a.forward and a.org_forward == orig
    b.apply_to()
    assert base.forward == b.forward and b.org_forward == a.forward
    assert base._lycoris_wrappers == [a, b]
    assert a._is_outermost_on_plain_forward() and not b._is_outermost_on_plain_forward()
    a.restore()  # remove the inner one: b must now sit on the original forward
    assert base.forward == b.forward and b.org_forward == orig
    assert b._is_outermost_on_plain_forward()
    b.restore()
    assert base.forward == orig
    assert not hasattr(base, "_lycoris_wrappers") and not hasattr(base, "_lycoris_original_forward")


def test_adapter_does_not_own_base_weights():
    base = nn.Linear(8, 8)
    m = L.LohaModule("a", base, 1.0, 2, 1)
    assert all("org_module" not in k for k in m.state_dict())
    assert sum(p.numel() for p in m.parameters()) == 4 * 8 * 2


def test_alpha_buffer_and_scale_rules():
    m = L.LoConModule("a", nn.Linear(8, 8), 1.0, 4, 0)  # alpha 0 -> alpha = dim
    assert m.scale == 1.0 and float(m.alpha) == 4.0
    m = L.LoConModule("a", nn.Linear(8, 8), 1.0, 4, 2, rs_lora=True)
    assert abs(m.scale - 1.0) < 1e-12 and abs(float(m.alpha) - 4.0) < 1e-6
    k = L.LokrModule("a", nn.Linear(64, 64), 1.0, 100000, 1, factor=8)  # both blocks full -> scale 1
    assert k.use_w1 and k.use_w2 and k.scale == 1.0
    assert tuple(k.lokr_w1.shape) == (8, 8) and tuple(k.lokr_w2.shape) == (8, 8)


def test_state_dict_scalar_folding_and_roundtrip(tmp_path):
    torch.manual_seed(0)
    unet = ToyUNet()
    net = create_network(1.0, 4, 1, None, None, unet, algo="lokr", preset="full", factor=4, conv_dim=4)
    net.apply_to(None, unet, False, True)
    with torch.no_grad():
        for p in net.parameters():
            p.add_(torch.randn_like(p) * 0.01)
    f = str(tmp_path / "net.safetensors")
    net.save_weights(f, torch.float32, {})
    from safetensors import safe_open

    with safe_open(f, "pt") as fh:
        meta = fh.metadata()
        keys = sorted(fh.keys())
    assert meta["sshs_model_hash"].startswith("0x") and len(meta["sshs_model_hash"]) == 66
    assert keys == sorted(net.state_dict().keys())
    net.restore()

    net2, sd = create_network_from_weights(1.0, f, None, None, unet)
    # order follows the checkpoint's key order (like the reference), so compare by name
    assert sorted(l.lora_name for l in net2.unet_loras) == sorted(l.lora_name for l in net.unet_loras)
    by_name = {l.lora_name: l for l in net2.unet_loras}
    for a in net.unet_loras:
        b = by_name[a.lora_name]
        assert type(a) is type(b)
        assert a.scale == b.scale
        for (ka, va), (kb, vb) in zip(a.state_dict().items(), b.state_dict().items()):
            if ka == "alpha" and a.use_w1 and a.use_w2:
                continue  # full x full LoKr re-derives lora_dim = 1 and forces alpha := 1 (lokr.py:262-269,209-211)
            assert ka == kb and torch.equal(va, vb), (a.lora_name, ka)


def test_generic_from_weights_roundtrip(tmp_path):
    torch.manual_seed(0)
    model = ToyUNet()
    for algo, kw in (("locon", {}), ("loha", {}), ("lokr", {"factor": 4})):
        LycorisNetwork.apply_preset({"target_module": ["Attention", "FeedForward"], "target_name": []})
        net = create_lycoris(model, 1.0, 4, 2, algo=algo, **kw)
        net.apply_to()
        with torch.no_grad():
            for p in net.parameters():
                p.add_(torch.randn_like(p) * 0.01)
        f = str(tmp_path / f"{algo}.pt")
        net.save_weights(f, None, None)
        net.restore()
        net2, _ = create_lycoris_from_weights(1.0, f, model)
        assert sorted(l.lora_name for l in net2.loras) == sorted(l.lora_name for l in net.loras)
        by_name = {l.lora_name: l for l in net2.loras}
        for a in net.loras:
            b = by_name[a.lora_name]
            for (ka, va), (kb, vb) in zip(a.state_dict().items(), b.state_dict().items()):
                assert ka == kb and torch.allclose(va, vb), (algo, a.lora_name, ka)


def test_prepare_optimizer_params_loraplus_groups():
    unet = ToyUNet()
    net = create_network(1.0, 4, 1, None, None, unet, algo="locon", preset="attn-mlp", loraplus_lr_ratio="4")
    groups, notes = net.prepare_optimizer_params(None, 1e-4, None)
    assert notes == ["unet", "unet plus"]
    assert groups[0]["lr"] == 1e-4 and groups[1]["lr"] == 4e-4
    n = sum(len(list(g["params"])) for g in groups)
    assert n == len(list(net.parameters())) or n == sum(len(list(l.parameters())) for l in net.loras)


def test_kohya_apply_to_requires_flags():
    unet = ToyUNet()
    net = create_network(1.0, 4, 1, None, None, unet, algo="locon", preset="attn-mlp")
    with pytest.raises(AssertionError):
        net.apply_to(None, unet)


def test_parametrize_api_registers():
    lin = nn.Linear(8, 8)
    L.LoConModule.parametrize(lin, "weight", 1.0, 2, 1)
    assert torch.nn.utils.parametrize.is_parametrized(lin, "weight")
    assert lin.weight.shape == (8, 8)  # CPU get_merged_weight is plain PyTorch (cold path)


def test_builtin_preset_tables_equal_the_reference():
    """lycoris/config.py PRESET (names, keys, target class / module-name lists incl. the DiT, Flux, SD3, Wan, Qwen
    and text-encoder class names) — the discovery lists that make other backbones work with no new math."""
    import json
    import os

    from conftest import GOLDEN
    from lycoris_b200.config import PRESET

    with open(os.path.join(GOLDEN, "presets.json")) as fh:
        ref = json.load(fh)
    assert json.loads(json.dumps(PRESET, sort_keys=True)) == ref


def test_plain_forward_detection_rejects_foreign_instance_patches():
    """Round-1 advisor finding: a base layer whose forward was already instance-patched (kohya networks.lora, an
    accelerate offload hook) must NOT be treated as pristine — the engine would contract org.weight directly and
    drop the other patch's contribution.  Only the class's own bound forward counts as plain."""
    import types

    import torch.nn as nn

    import lycoris_b200 as L

    lin = nn.Linear(16, 16)
    mod = L.LoConModule("a", lin, 1.0, 4, 1)
    mod.apply_to()
    assert mod._is_outermost_on_plain_forward()
    mod.restore()

    lin2 = nn.Linear(16, 16)
    orig = lin2.forward
    lin2.forward = lambda x: orig(x) + 1.0  # plain function
    m2 = L.LoConModule("b", lin2, 1.0, 4, 1)
    m2.apply_to()
    assert not m2._is_outermost_on_plain_forward()

    lin3 = nn.Linear(16, 16)
    lin3.forward = types.MethodType(lambda self, x: nn.Linear.forward(self, x) * 2, lin3)  # bound, but not the class's
    for cls, kw in ((L.LoConModule, {}), (L.IA3Module, {})):
        m3 = cls("c", lin3, 1.0, 4, 1, **kw)
        m3.apply_to()
        assert not m3._is_outermost_on_plain_forward(), cls.__name__
        m3.restore()

    # stacking: the second wrapper sits on the first wrapper's forward, not on the class forward
    lin4 = nn.Linear(16, 16)
    a, b = L.LoConModule("s1", lin4, 1.0, 4, 1), L.LoConModule("s2", lin4, 1.0, 4, 1)
    a.apply_to()
    b.apply_to()
    assert a._is_outermost_on_plain_forward() and not b._is_outermost_on_plain_forward()


def test_out_of_scope_checkpoint_entries_warn_instead_of_loading_silently(caplog):
    import logging

    import torch

    from lycoris_b200.modules import get_module

    sd = {"lora_unet_x.diff": torch.zeros(2, 2), "lora_unet_y.oft_blocks": torch.zeros(2, 2, 2),
          "lora_unet_z.lora_up.weight": torch.zeros(4, 2), "lora_unet_z.lora_down.weight": torch.zeros(2, 4)}
    logging.getLogger("LyCORIS").setLevel(logging.WARNING)
    with caplog.at_level(logging.WARNING, logger="LyCORIS"):
        assert get_module(sd, "lora_unet_x") == (None, None)
        assert get_module(sd, "lora_unet_y") == (None, None)
        cls, _ = get_module(sd, "lora_unet_z")
    logging.getLogger("LyCORIS").setLevel(logging.ERROR)
    assert cls.__name__ == "LoConModule"
    text = " ".join(r.getMessage() for r in caplog.records)
    assert "Full adapter" in text and "Diag-OFT/BOFT" in text


def test_quantised_style_base_is_forced_into_bypass_mode():
    """a base layer that is a Linear SUBCLASS (what bitsandbytes / quanto layers are — their `.weight` is
    not a dense 16-bit matrix) must never reach the merged-weight engine path: the adapter goes to bypass mode
    (reference base.py:162-177), whose forward only calls `org_forward` and the adapter's own small layers."""
    import torch
    import torch.nn as nn

    import lycoris_b200 as L

    class FakeQuantLinear(nn.Linear):  # dequantises on the fly; reading .weight directly would be wrong
        def forward(self, x):
            return nn.functional.linear(x, self.weight.detach().round(decimals=2), self.bias)

    torch.manual_seed(0)
    base = FakeQuantLinear(32, 48)
    base.requires_grad_(False)
    for cls, kw in ((L.LoConModule, {}), (L.LokrModule, {"factor": 4}), (L.LohaModule, {})):
        mod = cls("q", base, 1.0, 4, 2, **kw)
        assert mod.bypass_mode and mod.is_quant, cls.__name__
        with torch.no_grad():
            for p in mod.parameters():
                if float(p.abs().sum()) == 0.0:
                    p.normal_(0, 0.05)
        mod.apply_to()
        x = torch.randn(5, 32, requires_grad=True)
        y = base(x)  # CPU tensors: the rebuild-mode engine would raise EngineUnavailable; bypass is plain PyTorch
        want = FakeQuantLinear.forward(base, x) + mod.bypass_forward_diff(x, scale=mod.multiplier)
        assert torch.allclose(y, want, atol=1e-6), cls.__name__
        y.pow(2).mean().backward()
        assert all(p.grad is not None for p in mod.parameters()), cls.__name__
        mod.restore()
    # an explicit bypass_mode=False on a subclass is honoured (the user vouches for a dense weight)
    mod = L.LoConModule("q2", base, 1.0, 4, 2, bypass_mode=False)
    assert mod.bypass_mode is False and not mod.is_quant
