import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


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
