"""Detection of quantised base layers (bitsandbytes / quanto / optimum-quanto).

Such layers expose no dense 16-bit weight, so the adapter is forced into bypass mode
(reference lycoris/utils/quant.py:3-88, modules/base.py:162-177).  The B200 engine targets
bf16/fp16 base weights; quantised bases are recognised only to keep that rule.
"""

import importlib
from functools import lru_cache

from ..logging import logger

_CANDIDATES = (
    ("bitsandbytes.nn", ("Linear8bitLt", "LinearFP4", "LinearNF4")),
    ("quanto.nn", ("QLinear", "QConv2d", "QLayerNorm")),
    ("optimum.quanto.nn", ("QLinear", "QConv2d", "QLayerNorm")),
)


def _collect():
    found = []
    for mod_name, names in _CANDIDATES:
        try:
            mod = importlib.import_module(mod_name)
        except Exception:  # noqa: BLE001 - optional dependency, any import failure means "absent"
            continue
        found.extend(getattr(mod, n) for n in names if hasattr(mod, n))
    return tuple(found)


QuantLinears = _collect()
SUPPORT_QUANT = bool(QuantLinears)


@lru_cache(maxsize=None)
def log_bypass():
    return logger.warning("Using bnb/quanto/optimum-quanto with LyCORIS will enable force-bypass mode.")


@lru_cache(maxsize=None)
def log_suspect():
    return logger.warning(
        "Non-native Linear detected but bypass_mode is not set. "
        "Automatically using force-bypass mode to avoid possible issues. "
        "Please set bypass_mode=False explicitly if there are no quantized layers."
    )
