"""pytest configuration: `gpu` marker + shared helpers.

`-m "not gpu"` runs here (CPU only): oracle vs golden vectors, host logic, C-ABI symbol export.
`-m gpu` runs on the B200 box: parity of the CUDA path (through the C-ABI) against the oracle.
"""

from __future__ import annotations

import json
import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))

GOLDEN = ROOT / "tests" / "golden"


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def load_golden(name: str):
    return np.load(GOLDEN / name, allow_pickle=False)


def golden_json(npz, key: str):
    return json.loads(bytes(npz[key]).decode())


@pytest.fixture(scope="session")
def golden_dir() -> Path:
    return GOLDEN
