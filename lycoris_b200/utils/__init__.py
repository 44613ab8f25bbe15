"""Small helpers the preserved API calls.  The reference's offline tooling in this module
(SVD extraction, merge, key conversion — lycoris/utils/__init__.py:60-483) is out of scope."""

import hashlib
from io import BytesIO

from .general import product  # noqa: F401


def str_bool(val):
    """kohya passes network_args as strings: anything but "false" is true (utils/__init__.py:44)."""
    return str(val).lower() != "false"


def default(val, d):
    return d if val is None else val


def _tensor_payload(tensors):
    import safetensors.torch

    blob = BytesIO(safetensors.torch.save(tensors))
    header_len = int.from_bytes(blob.read(8), "little")
    blob.seek(8 + header_len)
    return blob.read()


def precalculate_safetensors_hashes(state_dict):
    """sshs model hash kohya stores in the metadata: sha256 over each tensor's safetensors payload,
    one tensor at a time (utils/__init__.py:19-41)."""
    digest = hashlib.sha256()
    for tensor in state_dict.values():
        digest.update(_tensor_payload({"tensor": tensor}))
    return f"0x{digest.hexdigest()}"
