"""pytest configuration: the ``gpu`` marker and shared fixtures.

``-m "not gpu"`` runs here on CPU (oracle vs golden vectors, host logic, C-ABI
symbol export); ``-m gpu`` runs on a B200 and is the parity suite proper.
"""
import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))

GOLDEN_DIR = Path(__file__).resolve().parent / "golden"
GOLDEN_CASES = sorted(p.stem for p in GOLDEN_DIR.glob("*.npz"))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def load_golden(name):
    with np.load(GOLDEN_DIR / f"{name}.npz") as z:
        return {k: z[k] for k in z.files}


@pytest.fixture(params=GOLDEN_CASES)
def golden(request):
    g = load_golden(request.param)
    g["name"] = request.param
    return g
