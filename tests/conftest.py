import json
import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent
GOLDEN = ROOT / "tests" / "golden"
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def rel_l2(a, b):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    nb = np.linalg.norm(b.ravel())
    return float(np.linalg.norm((a - b).ravel()) / (nb if nb > 0 else 1.0))


def rel_l2c(a, b):
    """rel-L2 for complex arrays."""
    a = np.asarray(a)
    b = np.asarray(b)
    nb = np.linalg.norm(b.ravel())
    return float(np.linalg.norm((a - b).ravel()) / (nb if nb > 0 else 1.0))


SOLVE_CASES = sorted(json.loads((GOLDEN / "index.json").read_text()).keys())


def load_case(name):
    """Golden case -> (solver kwargs without precision, dict of reference outputs)."""
    d = np.load(GOLDEN / f"solve_{name}.npz")
    meta = json.loads(str(d["meta"]))
    levels = d["levels"]
    if meta.pop("levels_scalar"):
        levels = int(levels)
    kw = dict(srf_flx=d["srf_flx"], z=d["z"],
              profiles=(d["u"], d["v"], d["Kx"], d["Ky"], d["Kz"]), levels=levels)
    for k, v in meta.items():
        kw[k] = tuple(v) if isinstance(v, list) else v
    return kw, d


@pytest.fixture(scope="session")
def oracle():
    from oracle import bldfm_oracle
    bldfm_oracle.build()
    return bldfm_oracle


@pytest.fixture(scope="session")
def gpu_lib():
    from bldfm_b200 import _lib
    if _lib.device_count() < 1:
        pytest.fail("no CUDA device visible: -m gpu tests must run on the GPU box")
    return _lib
