"""pytest configuration: `gpu` marker + shared helpers.

`-m "not gpu"` runs here (no GPU): oracle vs golden vectors, host logic, C-ABI symbol checks.
`-m gpu` runs on a B200: parity of the CUDA path (through the C-ABI) against the oracle/golden vectors.
"""
import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parents[1]
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))

GOLDEN = ROOT / "tests" / "golden"


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def golden():
    def load(name):
        return dict(np.load(GOLDEN / f"{name}.npz"))
    return load


def dense_grid(g, prefix, cells, dim):
    """Expand a sparsely stored golden grid (ids, v, m) to dense (cells,dim)/(cells,) arrays."""
    gv = np.zeros((cells, dim), np.float32)
    gm = np.zeros((cells,), np.float32)
    ids = g[f"{prefix}_ids"]
    gv[ids] = g[f"{prefix}_v"]
    gm[ids] = g[f"{prefix}_m"]
    return gv, gm
