"""pytest configuration: markers, import paths, shared fixtures.

`-m "not gpu"`: oracle vs golden fixtures, host logic, C-ABI surface (no GPU needed).
`-m gpu`     : parity of the CUDA path (through the C ABI) against the oracle and the fixtures.
"""
from __future__ import annotations

import sys
from pathlib import Path

import pytest
import torch

REPO = Path(__file__).resolve().parent.parent
GOLDEN = REPO / "tests" / "golden"
for p in (str(REPO), str(REPO / "tests" / "shim")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def pytest_collection_modifyitems(config, items):
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


def load_golden(name: str):
    return torch.load(GOLDEN / f"{name}.pt", weights_only=False)


@pytest.fixture(scope="session")
def golden():
    cache = {}

    def get(name: str):
        if name not in cache:
            cache[name] = load_golden(name)
        return cache[name]

    return get


@pytest.fixture(scope="session")
def sb():
    """The product package (comfyui-sonar_b200/), imported as `sonar_b200`."""
    import sonar_b200

    return sonar_b200


@pytest.fixture()
def cuda():
    return torch.device("cuda", 0)
