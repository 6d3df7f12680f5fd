"""pytest configuration: `gpu` marker, repo root on sys.path, golden-fixture loader."""

from __future__ import annotations

import sys
from pathlib import Path

import numpy as np
import pytest
import torch

ROOT = Path(__file__).resolve().parent.parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))

GOLDEN_DIR = Path(__file__).resolve().parent / "golden"


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: test needs a CUDA device (run with -m gpu on the B200 box)")


def pytest_collection_modifyitems(config, items):
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device in this container")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


class Golden:
    """Read-only view of one ``tests/golden/<name>.npz`` fixture with torch conversion."""

    def __init__(self, name: str):
        self._data = np.load(GOLDEN_DIR / f"{name}.npz", allow_pickle=False)

    def keys(self):
        return list(self._data.keys())

    def np(self, key: str) -> np.ndarray:
        return self._data[key]

    def t(self, key: str, device: str | torch.device = "cpu") -> torch.Tensor:
        return torch.from_numpy(np.ascontiguousarray(self._data[key])).to(device)


@pytest.fixture(scope="session")
def golden():
    cache: dict[str, Golden] = {}

    def load(name: str) -> Golden:
        if name not in cache:
            cache[name] = Golden(name)
        return cache[name]

    return load
