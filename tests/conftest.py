import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (sm_100a); run with -m gpu on the GPU box")


def _has_gpu() -> bool:
    try:
        import torch
        return torch.cuda.is_available() and torch.cuda.get_device_capability(0)[0] == 10
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    if _has_gpu():
        return
    skip = pytest.mark.skip(reason="no sm_100 GPU in this container")
    for it in items:
        if "gpu" in it.keywords:
            it.add_marker(skip)


@pytest.fixture(scope="session")
def lib():
    from qwen2_audio_whisper_ggml_b200 import lib as L
    return L.load_library()


@pytest.fixture(scope="session")
def ref():
    """the compiled reference (oracle/_ref); built here where /root/reference exists, shipped to the GPU box as a .so"""
    from oracle import refbind
    if not refbind.available():
        pytest.skip("oracle/_ref not built (make -C oracle ref)")
    refbind.load()
    return refbind
