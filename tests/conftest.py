import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with `pytest -m gpu` on the GPU box)")


def load_golden(name):
    z = np.load(os.path.join(GOLDEN, name + ".npz"))
    meta = json.loads(str(z["meta"])) if "meta" in z else None
    return z, meta


@pytest.fixture(scope="session")
def golden():
    return load_golden


_ckpt_cache = {}


def checkpoint(geometry_name, seed):
    """Regenerate the seeded random-init checkpoint a fixture was produced from (weights are not stored)."""
    import mgea_b200 as mg
    key = (geometry_name, seed)
    if key not in _ckpt_cache:
        _ckpt_cache[key] = mg.make_checkpoint(mg.GEOMETRIES[geometry_name], seed)
    return _ckpt_cache[key]


def has_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False
