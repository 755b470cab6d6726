import hashlib
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_collection_modifyitems(config, items):
    try:
        import torch
        have = torch.cuda.is_available()
    except Exception:  # pragma: no cover
        have = False
    if have:
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


def load_golden(name):
    return np.load(os.path.join(GOLDEN, name + ".npz"))


def sha(a) -> str:
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


@pytest.fixture(scope="session")
def golden():
    return load_golden


def tie_inputs(n=4096):
    """Exactly periodic, integer-valued inputs: p, 2p, 3p ... give bit-identical projections, so the reference's norms
    tie EXACTLY and its strict '>' keeps the lowest period (same generator as tests/golden/make_golden.py)."""
    idx = np.arange(n)
    out = {}
    for p in (3, 7, 10, 12, 25):
        rng = np.random.default_rng(900 + p)
        out[f"binary{p}"] = np.tile(rng.integers(0, 2, p).astype(float), n // p + 1)[:n]
        out[f"int{p}"] = np.tile(rng.integers(-5, 6, p).astype(float), n // p + 1)[:n]
    out["square8"] = np.sign(np.sin(2 * np.pi * (idx + 0.5) / 8))
    imp = np.zeros(n)
    imp[::9] = 1.0
    out["impulse9"] = imp
    return out


def refined_normal_equations(a, x, steps=8):
    """Solution of (A A^T) w = A x refined in extended precision (float64 LU as the preconditioner, residuals in
    np.longdouble): the yardstick against which the device's Cholesky and the reference's LU (np.linalg.solve,
    QOPeriods.py:794) are both measured.  Returns (w, residual x - A^T w) as longdouble, and the plain LU solution."""
    import scipy.linalg as sl
    gram = a @ a.T
    lu = sl.lu_factor(gram)
    w_lu = sl.lu_solve(lu, a @ x)
    gl, al = gram.astype(np.longdouble), a.astype(np.longdouble)
    bl = al @ x.astype(np.longdouble)
    w = w_lu.astype(np.longdouble)
    for _ in range(steps):
        w = w + sl.lu_solve(lu, (bl - gl @ w).astype(np.float64))
    return w, x.astype(np.longdouble) - al.T @ w, w_lu
