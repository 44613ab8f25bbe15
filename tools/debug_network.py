"""Locate a network-level mismatch: run the toy UNet through the engine and through the oracle patch, print the
relative error of out / dx / worst gradients, and the first wrapped layer (in backward order) whose input
gradient differs while its output gradient still agrees."""
import os
import sys

import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import test_gpu_network as T  # noqa: E402
from helpers import oracle_patch_network, rel_err  # noqa: E402

name, kw, preset, dim, regime = T.CASES[int(os.environ.get("CASE", "0"))]
unet, net, st = T._setup(kw, preset, dim, 2, True, regime)

rec = {}


def hook(tag, mname):
    def fn(mod, gin, gout):
        rec.setdefault(tag, []).append((mname, None if gin[0] is None else gin[0].detach().float().clone(),
                                        gout[0].detach().float().clone(),
                                        None if gin[0] is None else (tuple(gin[0].shape), gin[0].stride(), gin[0].dtype)))
    return fn


def run(tag):
    hs = [m.register_full_backward_hook(hook(tag, n)) for n, m in unet.named_modules()
          if isinstance(m, (torch.nn.Conv2d, torch.nn.Linear))]
    r = T._run(unet, net, st, regime)
    for h in hs:
        h.remove()
    return r


undo = oracle_patch_network(net)
ref = run("ref")
undo()
net.apply_to(None, unet, False, True)
got = run("eng")
net.restore()
errs = {k: rel_err(got[3][k], ref[3][k]) for k in ref[3] if float(ref[3][k].float().norm()) > 0}
worst = sorted(errs.items(), key=lambda kv: -kv[1])[:3]
print(name, "| out", round(rel_err(got[1], ref[1]), 4), "dx", round(rel_err(got[2], ref[2]), 4), "worst", worst)
for (n1, gi1, go1, _), (n2, gi2, go2, meta) in zip(rec["ref"], rec["eng"]):
    assert n1 == n2, (n1, n2)
    e_out = rel_err(go2, go1)
    e_in = None if gi1 is None or gi2 is None else rel_err(gi2, gi1)
    flag = "  <-- first bad input grad" if (e_in is not None and e_in > 0.1 and e_out < 0.1) else ""
    print(f"{n1:60s} dY err {e_out:.4f}  dX err {('%.4f' % e_in) if e_in is not None else 'n/a':>7s} {meta}{flag}")
    if flag:
        break
