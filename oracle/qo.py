"""Oracle for the QOPeriods default path (test infrastructure).

Numpy restatement of pyPeriod/QOPeriods.py:
  find_periods      :313-596 (default branch: _orthogonalize False, update_weights True)
  _update_weights   :598-643
  get_subspaces     :807-852
  Pp / Pp_column    :940-1003 (basis_type "natural"), Cq :1005-1052 (basis_type "ramanujan")
  solve_quadratic   :743-805
  get_periods       :719-741, concatenate_periods :854-887,
  stack_pairwise_gcd_subspaces :889-938, reduce_rows :86-94

Preserved reference behaviour: the default stop test compares rms(reconstruction)
(not the residual) with rms(data)*thresh (:391); when the test fails the weights
are re-solved with ALL periods but periods[:-1] are reported (:560-594); a
LinAlgError keeps the previous round's outputs (:552-559); a repeated best period
overwrites its dictionary entry with keep == 0, and `Pp(keep=0)` then returns all
p rows (`if keep:` at :972).  The stray `print(nonzero_periods)` at :488 is not
reproduced.
"""
from __future__ import annotations

import itertools

import numpy as np
from scipy import linalg as spla

from .numtheory import divisor_set, phi
from .periods import periodic_norm, project


def rms(x) -> float:
    """QOPeriods.py:78-79."""
    return np.sqrt(np.sum(np.power(x, 2)) / len(x))


def indicator_rows(p: int, n: int, keep=None) -> np.ndarray:
    """Natural-basis rows 1[(m - i) % p == 0], i < keep.  QOPeriods.py:940-1003."""
    p = int(p)
    m = np.arange(int(n))
    rows = ((m[None, :] - np.arange(p)[:, None]) % p == 0).astype(np.float64)
    if keep:  # keep == 0 or None keeps everything (:972)
        rows = rows[:keep]
    return rows


def ramanujan_rows(q: int, n: int, keep=None) -> np.ndarray:
    """Ramanujan-basis rows: row i = roll(c_q, i) tiled to n, c_q = real part of the sum of exp(2 pi i k m / q) over
    k coprime to q -- summed the way the reference sums it (QOPeriods.py:1037-1045), i.e. with its 1e-13 rounding
    noise; first `keep` rows (QOPeriods.py:968-974)."""
    q = int(q)
    vec = np.zeros(q, dtype=complex)
    ks = [k for k in range(1, q + 1) if np.gcd(k, q) == 1]
    for i in range(q):
        for k in ks:
            vec[i] = vec[i] + np.exp(1j * 2 * np.pi * k * i / q)
    reps = int(np.ceil(n / q))
    rows = np.zeros((q, int(n)))
    for i in range(q):
        rows[i] = np.real(np.tile(np.roll(vec, i), reps))[: int(n)]
    if keep:
        rows = rows[:keep]
    return rows


def get_subspaces(periods, n: int, basis: str = "natural"):
    """Stacked dictionary and {str(q): rows kept}.  QOPeriods.py:807-852."""
    rows_of = ramanujan_rows if basis == "ramanujan" else indicator_rows
    seen = set()
    dim_before = 0
    layout = {}
    for q in periods:
        seen = seen.union(divisor_set(int(q)))
        dim = int(np.sum([phi(r) for r in seen]))
        layout[str(q)] = dim - dim_before
        dim_before = dim
    blocks = [np.zeros((0, n))]
    for q, keep in layout.items():
        blocks.append(rows_of(int(q), n, keep))
    return np.vstack(blocks), layout


def get_subspaces_gcds_extracted(periods, n: int):
    """Dictionary of QOPeriodsWithGCDsExtracted (QOPeriodsWithGCDsExtracted.py:98-143): the found periods
    plus every common factor of every pair, iterated in the order of `set(sorted(P))`; a period keeps
    p - sum(phi(f)) rows over its proper factors f that are themselves in P (0 -> all rows, :972)."""
    q = [int(v) for v in periods]
    pset = set()
    for a, b in itertools.combinations(q, 2):
        pset = pset.union(set.intersection(divisor_set(a), divisor_set(b)))
    pset = pset.union(q)
    pset = set(sorted(pset))
    layout = {}
    blocks = [np.zeros((0, n))]
    for pp in pset:
        f = divisor_set(pp)
        if pp != 1:                      # get_factors(p, remove_n=True) keeps n when n == 1 (QOPeriods.py:73)
            f = f - {pp}
        f = f.intersection(pset)
        keep = int(pp - np.sum(np.array([phi(v) for v in f])))
        blocks.append(indicator_rows(pp, n, keep))
        layout[str(pp)] = keep
    return np.vstack(blocks), layout


def solve_quadratic(x: np.ndarray, a: np.ndarray, kind: str = "solve"):
    """Normal equations (A A^T) w = A x and reconstruction A^T w.  QOPeriods.py:743-805."""
    gram = np.matmul(a, a.T)
    rhs = np.matmul(a, x)
    if kind == "solve":
        w = np.linalg.solve(gram, rhs)
    else:
        w = np.linalg.lstsq(gram, rhs, rcond=None)[0]
    return w, np.matmul(a.T, w)


def find_periods(x: np.ndarray, num=None, thresh=None, min_length: int = 2, max_length=None,
                 trunc: bool = False, test_function=None, gcds_extracted: bool = False, basis: str = "natural"):
    """QOPeriods.find_periods, default branch.  QOPeriods.py:313-596.  `gcds_extracted` swaps in the
    dictionary layout of the QOPeriodsWithGCDsExtracted subclass (same loop, inherited); `basis` is the
    instance's basis_type ("natural" or "ramanujan", QOPeriods.py:153-155, 847-850)."""
    if gcds_extracted:
        get_subspaces = get_subspaces_gcds_extracted
    else:
        get_subspaces = lambda found, n_: globals()["get_subspaces"](found, n_, basis)
    n = len(x)
    if max_length is None:
        max_length = int(np.floor(n / 3))
    if num is None:
        num = n
    periods = np.zeros(num, dtype=np.uint32)
    norms = np.zeros(num)
    res = x.copy()
    if test_function is None:
        def test_function(_self, data, recon):  # :391
            return rms(recon) > (rms(data) * thresh)

    if np.sum(np.abs(x)) <= 1e-16:  # :394-406
        out = {"periods": np.array([1]), "norms": np.array([0]), "subspaces": np.ones((1, n)),
               "weights": np.array([0]), "basis_dictionary": {"1": n}}
        return out, np.zeros(n)
    out = {"periods": [], "norms": [], "subspaces": [], "weights": [], "basis_dictionary": {}}

    recon = None
    found = None
    for i in range(num):
        if i == 0 or test_function(None, x, recon):
            best_p, best = 0, 0
            for p in range(min_length, max_length + 1):  # :470-478
                val = periodic_norm(project(res, p, trunc, False), p)
                if val > best:
                    best_p, best = p, val
            periods[i] = best_p
            norms[i] = best
            found = periods[periods > 0]
            try:
                a, layout = get_subspaces(found, n)
                w, recon = solve_quadratic(x, a)  # against the ORIGINAL data (:517-522)
                res = x - recon
                out = {"periods": found, "norms": norms[: len(found)], "subspaces": a,
                       "weights": w, "basis_dictionary": layout}
            except np.linalg.LinAlgError:  # :552-559
                break
        else:  # :560-594
            a, layout = get_subspaces(found, n)
            w, recon = solve_quadratic(x, a)
            out = {"periods": found[:-1], "norms": norms[: len(found) - 1], "subspaces": a,
                   "weights": w, "basis_dictionary": layout}
            break
    return out, res


# --------------------------------------------------------------------------- get_periods
def concatenate_periods(weights, layout: dict) -> np.ndarray:
    """QOPeriods.py:854-887."""
    pos = 0
    parts = []
    for q, r in layout.items():
        v = np.zeros(int(q))
        v[0:r] = weights[pos: pos + r]
        pos += r
        parts.append(v)
    return np.array([t for part in parts for t in part])


def pairwise_gcd_rows(periods) -> np.ndarray:
    """QOPeriods.py:889-938."""
    periods = [int(p) for p in periods]
    if len(periods) > 1:
        rows = []
        for a, b in itertools.combinations(periods, 2):
            g = int(np.gcd(a, b))
            row = np.array([], dtype=np.int64)
            for p in periods:
                if p in (a, b):
                    comb = np.tile((np.arange(g) == 0).astype(np.float64), p // g)
                    seg = comb * -1 if p == a else comb
                else:
                    seg = np.zeros(p)
                row = np.append(row, seg)
            rows.append(row)
            for s in range(1, g):
                rows.append(np.roll(row, s))
        return np.vstack(tuple(rows))
    if len(periods) == 1:
        return np.ones((1, periods[0]))
    return np.ones((1, 1))


def reduce_rows(a: np.ndarray) -> np.ndarray:
    """Greedy rank-increasing row selection.  QOPeriods.py:86-94."""
    kept = a[0]
    rank = np.linalg.matrix_rank(kept)
    for row in a[1:]:
        trial = np.vstack((kept, row))
        if np.linalg.matrix_rank(trial) > int(rank):
            kept = trial
            rank = np.linalg.matrix_rank(trial)
    return kept


def get_periods(weights, layout: dict, decomp_type: str = "row reduction"):
    """QOPeriods.py:719-741."""
    periods = np.array([int(p) for p in layout.keys()])
    cat = concatenate_periods(weights, layout)
    a = pairwise_gcd_rows(periods)
    if decomp_type == "row reduction":
        _, rec = solve_quadratic(cat, reduce_rows(a), "solve")
    elif decomp_type == "lu":
        _, u = spla.lu(a, permute_l=True)
        _, rec = solve_quadratic(cat, u, "solve")
    elif decomp_type == "qr":
        _, r = np.linalg.qr(a, mode="complete")
        _, rec = solve_quadratic(cat, r, "solve")
    else:
        _, rec = solve_quadratic(cat, a, "lstsq")
    actual = cat - rec
    out = []
    for i, p in enumerate(periods):
        start = int(np.sum(periods[:i]))
        out.append(actual[start: start + p])
    return tuple(out)


# --------------------------------------------------------------------------- Muresan eq. 3 (QOPeriods.py:1122-1232)
def auto_corr(x: np.ndarray, k: int) -> float:
    """QOPeriods.py:1152-1173."""
    n = len(x)
    return float(np.sum(np.prod(np.vstack((x[0: n - k], x[k:n])), 0)))


def eq_3(x: np.ndarray, p: int) -> float:
    """QOPeriods.py:1123-1150: (P / N) * (ac(0) + 2 * sum_{l=1}^{M-1} ac(l P)), M = N // P."""
    n = len(x)
    second = 0
    for ell in range(1, n // p):
        second += auto_corr(x, int(ell * p))
    return (p / n) * (auto_corr(x, 0) + 2 * second)


def get_best_period_orthogonal(x: np.ndarray, max_p=None, normalize: bool = False, return_powers: bool = False):
    """QOPeriods.py:1175-1232."""
    if max_p is None:
        max_p = len(x) // 2
    q_all = np.arange(1, max_p)
    pows = np.zeros(q_all[-1] + 1)
    for q in q_all:
        pows[q] = max(eq_3(x, int(q)), 0)
        for f in divisor_set(int(q)):
            if f != q:
                pows[q] -= pows[f]
    pows[pows < 0] = 0
    if normalize:
        pows[1:] = pows[1:] / q_all
    if return_powers:
        return pows
    if np.argmax(pows) > 0:
        return q_all[np.argmax(pows)] - 1
    return 1
