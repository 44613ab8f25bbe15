import logging
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")
    logging.getLogger("LyCORIS").setLevel(logging.ERROR)


@pytest.fixture(autouse=True)
def _reset_class_config():
    """Presets mutate class attributes (reference quirk 9): restore them after every test."""
    from lycoris_b200.kohya import LycorisNetworkKohya
    from lycoris_b200.wrapper import LycorisNetwork

    saved = []
    for cls in (LycorisNetwork, LycorisNetworkKohya):
        saved.append((cls, {k: v for k, v in vars(cls).items() if k.isupper()}))
    yield
    for cls, attrs in saved:
        for k, v in attrs.items():
            setattr(cls, k, v)


def pytest_collection_modifyitems(config, items):
    """``gpu``-marked tests need a B200 and the built extension: skip them (instead of failing) on a
    box without CUDA, so a plain ``pytest tests`` is green on the CPU container too."""
    try:
        import torch

        have_gpu = torch.cuda.is_available()
    except Exception:  # noqa: BLE001
        have_gpu = False
    if have_gpu:
        return
    skip = pytest.mark.skip(reason="needs a CUDA device (B200); run with -m gpu on the GPU box")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


GOLDEN = os.path.join(ROOT, "tests", "golden")
