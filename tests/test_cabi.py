"""CPU: the C-ABI library loads, exports every symbol include/lyco_b200.h declares, and fails
loudly (no silent fallback) when there is no B200."""
import ctypes
import os
import re

import pytest
import torch

from conftest import ROOT
from lycoris_b200.engine import _lib


def _declared_symbols():
    src = open(os.path.join(ROOT, "include", "lyco_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(lyco_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_are_exported():
    lib = _lib.load()
    declared = _declared_symbols()
    assert declared, "no declarations parsed from include/lyco_b200.h"
    assert sorted(_lib.EXPORTED_SYMBOLS) == declared
    for sym in declared:
        assert hasattr(lib, sym), sym
    assert lib.lyco_abi_version() == 3


def test_delta_desc_layout_matches_header():
    # 14 int32 + 4 float + 4 pointers, naturally aligned
    assert ctypes.sizeof(_lib.DeltaDesc) == 14 * 4 + 4 * 4 + 4 * 8
    names = [f[0] for f in _lib.DeltaDesc._fields_]
    hdr = open(os.path.join(ROOT, "include", "lyco_b200.h")).read()
    body = hdr[hdr.index("typedef struct lyco_delta_desc"):hdr.index("} lyco_delta_desc_t;")]
    body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
    decl = []
    for line in body.splitlines():
        m = re.match(r"\s*(?:const\s+)?(?:int32_t|float|void\*)\s+([^;]+);", line)
        if m:
            decl += [n.strip() for n in m.group(1).split(",")]
    assert decl == names


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU failure mode")
def test_no_gpu_means_loud_failure():
    lib = _lib.load()
    assert lib.lyco_device_check(0) != 0
    assert len(_lib.last_error()) > 0


def test_cpu_forward_raises_instead_of_falling_back():
    import torch.nn as nn

    from lycoris_b200.engine._lib import EngineUnavailable
    from lycoris_b200.modules import LoConModule

    base = nn.Linear(16, 16)
    mod = LoConModule("t", base, 1.0, 4, 1)
    mod.apply_to()
    with pytest.raises(EngineUnavailable):
        base(torch.randn(2, 16))
    mod.restore()
    assert base(torch.randn(2, 16)).shape == (2, 16)


def test_no_product_import_of_oracle():
    """The product package must never import the oracle (or any CPU fallback)."""
    pkg = os.path.join(ROOT, "lycoris_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(".py"):
                src = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in src.replace("oracle/", ""), os.path.join(dirpath, f)
