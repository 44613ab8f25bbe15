"""ctypes binding of the C-ABI in ``include/lyco_b200.h``.

The shared library is built in-tree by ``__graft_entry__.build()`` (plain ``nvcc -shared``) as
``lycoris_b200/_lyco_b200.so``.  There is deliberately no CPU implementation behind it: if the
library is missing, or the device is not a B200, every op raises.
"""

from __future__ import annotations

import ctypes
import os
from ctypes import POINTER, Structure, c_char_p, c_float, c_int, c_int32, c_int64, c_uint64, c_void_p

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(os.path.dirname(_HERE), "_lyco_b200.so")

BF16, F16, F32 = 0, 1, 2
ALGO_LOCON, ALGO_LOHA, ALGO_LOKR, ALGO_IA3, ALGO_DYLORA, ALGO_RAW = 0, 1, 2, 3, 4, 5

# every symbol include/lyco_b200.h declares (checked by the CPU test-suite)
EXPORTED_SYMBOLS = (
    "lyco_abi_version",
    "lyco_last_error",
    "lyco_device_check",
    "lyco_launch_count",
    "lyco_gemm",
    "lyco_gemm_dual",
    "lyco_conv2d_fprop",
    "lyco_conv2d_wgrad",
    "lyco_transpose_cast",
    "lyco_filter_relayout",
    "lyco_merge_weight",
    "lyco_factor_grads",
    "lyco_grad_prep",
    "lyco_hada",
    "lyco_lokr_mix",
    "lyco_lokr_w1grad",
    "lyco_delta_weight",
    "lyco_dora_fwd",
    "lyco_dora_bwd",
)
ABI_VERSION = 3


class EngineUnavailable(RuntimeError):
    """The CUDA extension is missing or unusable; there is no fallback path."""


class DeltaDesc(Structure):
    """Mirror of ``lyco_delta_desc_t``."""

    _fields_ = [
        ("algo", c_int32),
        ("out_dim", c_int32),
        ("in_dim", c_int32),
        ("rank", c_int32),
        ("up", c_int32),
        ("uq", c_int32),
        ("vp", c_int32),
        ("vq", c_int32),
        ("on_input", c_int32),
        ("ia3_group", c_int32),
        ("f_dtype", c_int32),
        ("w_dtype", c_int32),
        ("pre_round", c_int32),
        ("pre_dtype", c_int32),
        ("m_in", c_float),
        ("m_pre", c_float),
        ("m_post1", c_float),
        ("m_post2", c_float),
        ("f0", c_void_p),
        ("f1", c_void_p),
        ("f2", c_void_p),
        ("f3", c_void_p),
    ]


_lib = None


def _bind(lib):
    lib.lyco_abi_version.restype = c_int
    lib.lyco_abi_version.argtypes = []
    lib.lyco_last_error.restype = c_char_p
    lib.lyco_last_error.argtypes = []
    lib.lyco_device_check.restype = c_int
    lib.lyco_device_check.argtypes = [c_int]
    lib.lyco_launch_count.restype = c_uint64
    lib.lyco_launch_count.argtypes = []
    lib.lyco_gemm.restype = c_int
    lib.lyco_gemm.argtypes = [
        c_void_p, c_int, c_int64,  # A
        c_void_p, c_int, c_int64,  # B
        c_void_p, c_int, c_int64,  # C
        c_void_p, c_int,  # bias
        c_int, c_int, c_int,  # M N K
        c_int, c_int, c_int, c_void_p,  # ab_dtype, split_k, accumulate, stream
    ]
    lib.lyco_gemm_dual.restype = c_int
    lib.lyco_gemm_dual.argtypes = [
        c_void_p, c_int64, c_void_p, c_int, c_int64, c_int,  # A lda B b_mn ldb K
        c_void_p, c_int64, c_void_p, c_int64, c_int,  # A2 lda2 B2 ldb2 K2
        c_void_p, c_int64, c_void_p, c_int, c_int, c_int, c_int, c_void_p,  # C ldc bias bias_dtype M N dtype stream
    ]
    lib.lyco_conv2d_fprop.restype = c_int
    lib.lyco_conv2d_fprop.argtypes = [
        c_void_p, c_void_p, c_void_p, c_void_p, c_int,  # X Wk Y bias bias_dtype
        c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_int,  # Nb H W C O R S pad_h pad_w stride
        c_int, c_int, c_void_p,  # dtype y_layout stream
    ]
    lib.lyco_filter_relayout.restype = c_int
    lib.lyco_filter_relayout.argtypes = [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p]
    lib.lyco_transpose_cast.restype = c_int
    lib.lyco_transpose_cast.argtypes = [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p]
    lib.lyco_conv2d_wgrad.restype = c_int
    lib.lyco_conv2d_wgrad.argtypes = [
        c_void_p, c_void_p, c_void_p,  # X dY dW
        c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_int,  # Nb H W C O R S pad_h pad_w stride
        c_int, c_int, c_void_p,  # dtype split_k stream
    ]
    lib.lyco_merge_weight.restype = c_int
    lib.lyco_merge_weight.argtypes = [POINTER(DeltaDesc), c_void_p, c_void_p, c_void_p]
    lib.lyco_factor_grads.restype = c_int
    lib.lyco_factor_grads.argtypes = [
        POINTER(DeltaDesc), c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
    ]
    lib.lyco_grad_prep.restype = c_int
    lib.lyco_grad_prep.argtypes = [c_void_p, c_void_p, c_void_p, c_int64, c_float, c_int, c_void_p]
    lib.lyco_hada.restype = c_int
    lib.lyco_hada.argtypes = [
        c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,  # mode w1a w1b w2a w2b W out0 out1
        c_int, c_int, c_int, c_int, c_int, c_float, c_float, c_float, c_float, c_void_p,
    ]
    lib.lyco_lokr_mix.restype = c_int
    lib.lyco_lokr_mix.argtypes = [
        c_void_p, c_void_p, c_void_p, c_int, c_int, c_int,  # in out w w_dtype ldw transpose
        c_int64, c_int, c_int, c_int, c_int, c_void_p, c_int64, c_void_p,  # M na nb nc dtype zero_buf zero_n stream
    ]
    lib.lyco_lokr_w1grad.restype = c_int
    lib.lyco_lokr_w1grad.argtypes = [
        c_void_p, c_void_p, c_void_p, c_int64, c_int, c_int, c_int, c_float, c_int, c_int, c_void_p,
    ]
    lib.lyco_delta_weight.restype = c_int
    lib.lyco_delta_weight.argtypes = [POINTER(DeltaDesc), c_void_p, c_void_p, c_int, c_void_p, c_void_p]
    lib.lyco_dora_fwd.restype = c_int
    lib.lyco_dora_fwd.argtypes = [
        c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_float, c_float, c_int, c_int, c_void_p,
    ]
    lib.lyco_dora_bwd.restype = c_int
    lib.lyco_dora_bwd.argtypes = [
        c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,  # dW Wm dora_scale sumsq t g_scale
        c_int, c_int, c_int, c_int, c_float, c_float, c_int, c_int, c_void_p,
    ]
    return lib


def load():
    """Load (once) and return the bound library; raise :class:`EngineUnavailable` if absent."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise EngineUnavailable(
            f"lycoris_b200: CUDA extension not built ({LIB_PATH} missing). "
            "Run `python -c 'import __graft_entry__ as g; g.build()'` from the repo root. "
            "There is no CPU or PyTorch fallback for the adapter hot path."
        )
    try:
        lib = ctypes.CDLL(LIB_PATH)
    except OSError as e:  # pragma: no cover - depends on the box
        raise EngineUnavailable(f"lycoris_b200: cannot load {LIB_PATH}: {e}") from e
    _lib = _bind(lib)
    if _lib.lyco_abi_version() != ABI_VERSION:
        raise EngineUnavailable("lycoris_b200: ABI version mismatch between _lib.py and the .so")
    return _lib


def last_error() -> str:
    return load().lyco_last_error().decode("utf-8", "replace")


def check(rc: int, what: str):
    if rc != 0:
        raise RuntimeError(f"lycoris_b200 {what} failed: {last_error()}")


def launch_count() -> int:
    return int(load().lyco_launch_count())
