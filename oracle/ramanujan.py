"""Oracle for RamanujanPeriods (test infrastructure).

Numpy restatement of pyPeriod/RamanujanPeriods.py:
  Cq                          :133-154  (brute-force complex exponential sum)
  Cq_complete                 :156-169  (q rows -- the totient line :159 is commented out)
  project                     :124-131  (float32 storage, :127)
  find_periods                :67-86
  find_periods_with_weights   :88-122

`ramanujan_sum_exact` / `norms_closed_form_f64` are NOT reference code: they are
the integer-exact / fp64 closed forms of SURVEY.md §8a rows 7-8 that the CUDA
path implements, kept here so tests can separate "float32 storage noise of the
reference" (5e-9..2e-7 relative) from real disagreement.
"""
from __future__ import annotations

from math import gcd

import numpy as np

from .numtheory import phi
from .qo import get_subspaces, solve_quadratic


def cq(q: int) -> np.ndarray:
    """Real Ramanujan sum c_q(n), n < q, by the reference's complex sum.  :133-148."""
    q = int(q)
    ks = [k for k in range(1, q + 1) if gcd(k, q) == 1]
    vec = np.zeros(q, dtype=complex)
    for i in range(q):
        for k in ks:
            vec[i] = vec[i] + np.exp(1j * 2 * np.pi * k * i / q)
    return np.real(vec)


def cq_complete(q: int, n: int, normalize: bool = True) -> np.ndarray:
    """(q, n) dictionary of rolled, tiled, L2-normalised c_q rows.  :156-169."""
    base = cq(q)
    reps = int(np.ceil(n / q))
    mat = np.zeros((q, int(n)))
    for i in range(q):
        mat[i] = np.tile(np.roll(base, i), reps)[:n]
        if normalize:
            mat[i] /= np.linalg.norm(mat[i])
    return mat


def project_f32(x: np.ndarray, basis: np.ndarray) -> np.ndarray:
    """Row-wise <x,row>*row with row rescaled by its max, stored in float32.  :124-131."""
    out = np.zeros(basis.shape, dtype=np.float32)
    for i, row in enumerate(basis):
        row = row / np.max(row)
        out[i] = np.dot(x, row) * row
    return out


def find_periods(x: np.ndarray, min_length: int = 2, max_length=None) -> np.ndarray:
    """Ramanujan periodogram, length max_length+1.  :67-86."""
    if not max_length:
        max_length = len(x) // 3
    norms = np.zeros(max_length + 1)
    for q in range(min_length, max_length + 1):
        total = np.sum(project_f32(x, cq_complete(q, len(x))), 0)
        norms[q] = np.sum(np.power(total, 2))
    return norms


def find_periods_with_weights(x: np.ndarray, min_length: int = 2, max_length=None,
                              thresh: float = 0.2, norms: np.ndarray | None = None):
    """:88-122 (with the P7 unpack-order fix).  `norms` may be supplied to skip the slow sweep."""
    if norms is None:
        norms = find_periods(x, min_length, max_length)
    periods = np.argwhere(norms / np.abs(np.max(norms)) > thresh).flatten()
    a, layout = get_subspaces(periods, len(x))
    w, recon = solve_quadratic(x, a)
    res = x - recon
    return ({"periods": periods, "norms": norms[periods], "subspaces": a, "weights": w,
             "basis_dictionary": layout}, res)


def find_periods_f32_exact_cq(x: np.ndarray, min_length: int = 2, max_length=None) -> np.ndarray:
    """find_periods (:67-86) with project's float32 storage (:124-131), the rows built from the integer-exact c_q
    instead of the complex sum (:133-148), which is the only difference from `find_periods` above (and ~1000x
    faster).  Bit-identical to the reference on the committed fixtures (tests/test_oracle_golden.py); pins the
    CUDA path's precision="f32_compat" mode."""
    n = len(x)
    if not max_length:
        max_length = n // 3
    norms = np.zeros(max_length + 1)
    idx = np.arange(n)
    for q in range(min_length, max_length + 1):
        r = ramanujan_sum_exact(q) / phi(q)                       # row / max(row)
        rows = r[(idx[None, :] - np.arange(q)[:, None]) % q]      # (q, n): roll by i, tiled
        out = np.zeros((q, n), dtype=np.float32)
        for i in range(q):
            out[i] = np.dot(x, rows[i]) * rows[i]
        total = np.sum(out, 0)
        norms[q] = np.sum(np.power(total, 2))
    return norms


# ------------------------------------------------------------------ closed forms (not reference code)
def ramanujan_sum_exact(q: int) -> np.ndarray:
    """c_q(n) = mu(q/g) * phi(q) / phi(q/g), g = gcd(n, q); integer-valued."""
    q = int(q)

    def mobius(m: int) -> int:
        res, d = 1, 2
        while d * d <= m:
            if m % d == 0:
                m //= d
                if m % d == 0:
                    return 0
                res = -res
            d += 1
        return -res if m > 1 else res

    out = np.zeros(q)
    for n in range(q):
        g = gcd(n, q) if n else q
        out[n] = mobius(q // g) * phi(q) // phi(q // g)
    return out


def norms_closed_form_f64(x: np.ndarray, min_length: int = 2, max_length=None) -> np.ndarray:
    """fp64 closed form of the periodogram: y = C S_q, z = C^T y, norm = sum cnt*z^2."""
    n = len(x)
    if not max_length:
        max_length = n // 3
    norms = np.zeros(max_length + 1)
    idx = np.arange(n)
    for q in range(min_length, max_length + 1):
        c = ramanujan_sum_exact(q) / phi(q)
        s = np.bincount(idx % q, weights=x, minlength=q)
        cnt = np.bincount(idx % q, minlength=q)
        mat = c[(np.arange(q)[None, :] - np.arange(q)[:, None]) % q]  # C[i, m] = c((m - i) mod q)
        z = mat.T @ (mat @ s)
        norms[q] = np.sum(cnt * z * z)
    return norms
