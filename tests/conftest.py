import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def _gpu_usable():
    try:
        import torch
        if not torch.cuda.is_available():
            return False, "no CUDA device"
    except Exception as exc:          # torch missing
        return False, "torch unavailable: %s" % exc
    lib = os.path.join(ROOT, "ship_sim_gym_b200", "libshipsim.so")
    if not os.path.exists(os.environ.get("SHIPSIM_LIB") or lib):
        # on a GPU box a missing extension must FAIL loudly, not skip: the product has no fallback
        return True, ""
    return True, ""


def pytest_collection_modifyitems(config, items):
    """A plain `pytest tests` on a CPU-only box skips the gpu-marked tests instead of dying in torch._C._cuda_init."""
    ok, why = _gpu_usable()
    if ok:
        return
    skip = pytest.mark.skip(reason="gpu test: " + why)
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def golden_dir():
    return os.path.join(ROOT, "tests", "golden")
