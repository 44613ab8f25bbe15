"""Adapter modules on the hot path (LoCon, LoHa, LoKr, (IA)^3, DyLoRA).

The reference also ships Full / Norm / GLoRA / Diag-OFT / BOFT modules
(lycoris/modules/__init__.py:19-30); they are outside the scope table and are not
provided — asking for them raises a clear error in the wrapper.
"""

import torch

from ..functional.general import factorization  # noqa: F401
from .base import LycorisBaseModule
from .dylora import DyLoraModule
from .ia3 import IA3Module
from .locon import LoConModule
from .loha import LohaModule
from .lokr import LokrModule

# detection order matters: first class whose marker key is present wins (modules/__init__.py:19-37)
MODULE_LIST = [LoConModule, LohaModule, IA3Module, LokrModule, DyLoraModule]


# marker keys of the reference's adapter classes that are outside the scope table (full.py:28, norms.py:15,
# glora.py:27, diag_oft.py:34 / boft.py:49): a checkpoint carrying them must not load as a silently partial network
_OUT_OF_SCOPE_MARKERS = {"diff": "Full", "w_norm": "Norm", "a1.weight": "GLoRA", "oft_blocks": "Diag-OFT/BOFT"}


def get_module(lyco_state_dict, lora_name):
    """(adapter class, its tensors in ``weight_list`` order) for ``lora_name`` in a checkpoint."""
    for module in MODULE_LIST:
        if module.algo_check(lyco_state_dict, lora_name):
            return module, tuple(module.extract_state_dict(lyco_state_dict, lora_name))
    for key, algo in _OUT_OF_SCOPE_MARKERS.items():
        if f"{lora_name}.{key}" in lyco_state_dict:
            from ..logging import logger

            logger.warning(
                f"lycoris_b200: checkpoint entry {lora_name!r} is a {algo} adapter, which this engine does not "
                "provide (SURVEY.md section 8 scope); it is SKIPPED and the loaded network is partial")
            break
    return None, None


@torch.no_grad()
def make_module(lyco_type: LycorisBaseModule, params, lora_name, orig_module):
    try:
        return lyco_type.make_module_from_state_dict(lora_name, orig_module, *params)
    except NotImplementedError:
        return None
