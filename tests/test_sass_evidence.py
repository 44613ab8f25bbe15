"""The built library must contain the Blackwell instructions the design claims (checked on the SASS, no GPU needed):
tcgen05.mma as UTCHMMA (single CTA and .2CTA pairs), tcgen05.ld as LDTM, TMA loads / im2col gathers / stores as
UTMALDG / UTMASTG, the TMA reduce-add of the fp32 epilogue as UTMAREDG, programmatic dependent launch as
PREEXIT (griddepcontrol.launch_dependents) / ACQBULK (griddepcontrol.wait) — and no legacy HMMA tensor-core path."""
import os
import re
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SO = os.path.join(ROOT, "lycoris_b200", "_lyco_b200.so")


@pytest.fixture(scope="module")
def sass():
    tool = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(tool):
        pytest.skip("cuobjdump not available")
    if not os.path.exists(SO):
        pytest.skip("extension not built (run __graft_entry__.build())")
    out = subprocess.run([tool, "-sass", SO], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr[-500:]
    return out.stdout


def _count(sass, pattern):
    return len(re.findall(pattern, sass))


def test_built_for_sm100a_only(sass):
    archs = set(re.findall(r"arch = (sm_\w+)", sass))
    assert archs == {"sm_100a"}, archs


@pytest.mark.parametrize("mnemonic, why", [
    (r"\bUTCHMMA\b", "tcgen05.mma"),
    (r"\bUTCHMMA\.2CTA\b", "tcgen05.mma.cta_group::2 (CTA pairs)"),
    (r"\bLDTM\b", "tcgen05.ld (TMEM -> registers)"),
    (r"\bUTMALDG\.2D\b", "TMA tile loads"),
    (r"\bUTMALDG\.2D\.2CTA\b", "2-CTA TMA loads signalling the leader's mbarrier"),
    (r"\bUTMALDG\.4D\.IM2COL\b", "TMA im2col gathers of the convolution kernels"),
    (r"\bUTMASTG\.2D\b", "TMA stores of the epilogue"),
    (r"\bUTMAREDG\.2D\.ADD\b", "cp.reduce.async.bulk.tensor .add (fp32 split-K epilogue)"),
    (r"\bUTCBAR\.2CTA\.MULTICAST\b", "tcgen05.commit multicast to both CTAs of a pair"),
    (r"\bPREEXIT\b", "griddepcontrol.launch_dependents"),
    (r"\bACQBULK\b", "griddepcontrol.wait"),
])
def test_mnemonic_present(sass, mnemonic, why):
    assert _count(sass, mnemonic) > 0, f"no {mnemonic} in the SASS: {why} is not in the built library"


def test_no_legacy_tensor_core_path(sass):
    # mma.sync / wmma compile to HMMA: the recompiled-for-sm_100a baseline this engine is meant to replace
    assert _count(sass, r"\bHMMA\b") == 0


def test_tensor_core_kernels_are_the_hot_kernels(sass):
    # every gemm / conv / hada instantiation issues UTCHMMA; the HBM-bound helpers must not
    funcs = re.split(r"\n\s*Function : ", sass)
    with_mma = [f.split("\n", 1)[0] for f in funcs[1:] if "UTCHMMA" in f]
    assert with_mma, "no kernel issues tcgen05.mma"
    assert all(("gemm_sm100" in n) or ("conv_sm100" in n) or ("hada_sm100" in n) for n in with_mma), with_mma
