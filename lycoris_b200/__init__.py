"""lycoris_b200 — B200-native adapter-layer engine behind the LyCORIS API.

Drop-in for the forward/backward hot path of KohakuBlueleaf/LyCORIS (LoCon / LoHa / LoKr /
(IA)^3 / DyLoRA on ``nn.Linear`` / ``nn.Conv2d``): same ``wrapper`` / ``kohya`` / ``modules`` API,
hand-written sm_100a CUDA kernels (tcgen05 + TMA) underneath, reached through the C-ABI in
``include/lyco_b200.h``.  See DESIGN.md / INTEGRATION.md.
"""

try:
    from . import kohya
except Exception:  # noqa: BLE001 - same tolerance as the reference package init
    pass
from . import modules, utils
from .logging import logger
from .modules import make_module
from .modules.dylora import DyLoraModule
from .modules.ia3 import IA3Module
from .modules.locon import LoConModule
from .modules.loha import LohaModule
from .modules.lokr import LokrModule
from .wrapper import LycorisNetwork, create_lycoris, create_lycoris_from_weights

__version__ = "0.1.0"
