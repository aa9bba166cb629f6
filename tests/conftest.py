import os
import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))

GOLDEN = Path(__file__).resolve().parent / "golden"


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA (B200, sm_100a) device")


@pytest.fixture(scope="session")
def golden():
    return dict(np.load(GOLDEN / "reference_small.npz"))


@pytest.fixture(scope="session")
def golden_b32():
    return dict(np.load(GOLDEN / "b32_config1.npz"))


def relerr(a, b):
    a = np.asarray(a, np.float64)
    b = np.asarray(b, np.float64)
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300))
