"""Round-2 golden fixtures (run in the build container: ``python oracle/gen_golden_tucker.py``; needs /root/reference).

SURVEY.md §8 rows a8 / f3: the Tucker variants the round-1 fixtures did not pin — LoHa-Tucker
(lycoris/functional/loha.py:33-75 HadaWeightTucker) and LoKr-Tucker (lycoris/modules/lokr.py:121-128, 362-366) —
plus Tucker and DoRA cases on a 64-channel convolution, the smallest shape the engine's TMA-im2col kernels take, so
that an option-variant layer reaches conv_sm100_kernel instead of the library fallback.

Same procedure as oracle/gen_golden.py (whose ``run_case`` is reused): the unmodified reference module runs forward +
backward on the CPU, the oracle runs on the same tensors and must reproduce it BIT-EXACTLY, and the reference's
outputs are stored as tests/golden/tucker_{regime}.pt.
"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import gen_golden as G  # noqa: E402  (imports the real reference from /root/reference)

G.ALGOS["loha_tucker"] = dict(cls="LohaModule", dim=4, alpha=2.0, kw={})
G.ALGOS["lokr_tucker"] = dict(cls="LokrModule", dim=2, alpha=1.0, kw={"factor": 4})

TUCKER_CASES = {
    "loha_tucker/conv3": dict(algo="loha_tucker", layer="conv3", kw={}, use_tucker=True),
    "lokr_tucker/conv3": dict(algo="lokr_tucker", layer="conv3", kw={}, use_tucker=True),
    "locon_tucker/conv3c64": dict(algo="locon", layer="conv3c64", kw={}, use_tucker=True),
    "loha_tucker/conv3c64": dict(algo="loha_tucker", layer="conv3c64", kw={}, use_tucker=True),
    "lokr_tucker/conv3c64": dict(algo="lokr_tucker", layer="conv3c64", kw={}, use_tucker=True),
    "locon_dora/conv3c64": dict(algo="locon", layer="conv3c64", kw={"weight_decompose": True}),
    "loha_dora_in/conv3c64": dict(algo="loha", layer="conv3c64", kw={"weight_decompose": True, "wd_on_out": False}),
    "lokr_dora/conv3c64": dict(algo="lokr_full", layer="conv3c64", kw={"weight_decompose": True}),
    "lokr_full/conv3c64": dict(algo="lokr_full", layer="conv3c64", kw={}),
}


def main():
    n = 0
    for regime in G.REGIMES:
        out = {}
        seed = 5000
        for name, c in TUCKER_CASES.items():
            seed += 10
            case = G.run_case(c["algo"], c["layer"], regime, seed, c["kw"], 1.0, c.get("use_tucker", False))
            if c.get("use_tucker"):
                keys = set(case["params"])
                assert keys & {"lora_mid.weight", "hada_t1", "lokr_t2"}, (name, sorted(keys))  # a Tucker core exists
            out[name] = case
            n += 1
        torch.save(out, os.path.join(G.OUT, f"tucker_{regime}.pt"))
    print(f"{n} Tucker / 64-channel option cases: oracle == reference (bit-exact), fixtures written")


if __name__ == "__main__":
    main()
