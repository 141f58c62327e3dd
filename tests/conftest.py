import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def _gpu_available():
    try:
        from victor_b200 import _lib
        return _lib.load().vb200_device_count() > 0
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    """A plain `pytest tests` on a box without a CUDA device skips the GPU tests instead of failing them
    (the product itself never falls back: tests/test_host_tables.py::test_no_cpu_fallback_without_a_gpu)."""
    gpu_items = [it for it in items if "gpu" in it.keywords]
    if not gpu_items or _gpu_available():
        return
    skip = pytest.mark.skip(reason="needs a CUDA device and the built libvictor_b200.so")
    for it in gpu_items:
        it.add_marker(skip)


@pytest.fixture(scope="session")
def repo_root():
    return ROOT


@pytest.fixture(scope="session")
def boss_blocks():
    import yaml
    with open(os.path.join(ROOT, "config", "boss_config.yaml")) as fh:
        info = yaml.full_load(fh)
    info["model"]["dir"] = ROOT
    info["data"]["dir"] = ROOT
    return info["model"], info["data"]


@pytest.fixture(scope="session")
def example_block():
    import yaml
    with open(os.path.join(ROOT, "config", "example_model_input.yaml")) as fh:
        model = yaml.full_load(fh)["model"]
    model["dir"] = ROOT
    return model


def load_golden(name):
    import numpy as np
    return np.load(os.path.join(GOLDEN, name + ".npz"))


@pytest.fixture(scope="session")
def golden():
    return load_golden
