"""Oracle for the Sethares-Staley `Periods` class (test infrastructure).

Numpy restatement of pyPeriod/Periods.py, arithmetic kept operation-for-operation
where it decides bits:
  project            Periods.py:142-219
  periodic_norm      Periods.py:221-241
  small_to_large     Periods.py:246-287
  best_correlation   Periods.py:289-349
  m_best_meta        Periods.py:456-601 (via m_best :408-430, m_best_gamma :432-454)

Deliberately preserved reference behaviour: `orthogonalize` is
honoured inside M-best although a warning says otherwise (:482-483, :504-506);
the gamma norm in M-best step 2 divides by sqrt(max_length), the stale step-1
loop variable (:559, :572); `nq` is the norm of the LAST factor's projection,
not of xq (:570-572); the step-2 `while changed` runs exactly once (:541-544).
The only intended difference from the reference is speed: the per-residue
divisor vector is built with slicing instead of a Python loop (:188-193); the
values and the order of every floating-point operation are unchanged.
"""
from __future__ import annotations

import math
import warnings

import numpy as np

from .numtheory import factors_in_set_order, orth_chain


def _fold_rect(x: np.ndarray, p: int):
    """Zero-pad x to a multiple of p and view it as (rows, p).  Periods.py:171-176."""
    n = x.size
    short = int(np.ceil(n / p) * p - n)
    rect = np.zeros(n + short, dtype=x.dtype)
    rect[:n] = x
    return rect.reshape((n + short) // p, p), short


def project(x: np.ndarray, p: int = 2, trunc: bool = False, orth: bool = False,
            single_period: bool = False) -> np.ndarray:
    """Projection of x onto the p-periodic subspace.  Periods.py:142-219."""
    p = int(p)
    rect, short = _fold_rect(x, p)
    rows = rect.shape[0]
    if trunc:
        # Periods.py:178-184: mean over the complete rows only
        one = np.mean(rect, 0) if short == 0 else np.mean(rect[:-1], 0)
    else:
        # Periods.py:185-194: column sums over the padded rectangle / per-column counts
        counts = np.full(p, float(rows))
        counts[p - short:] = rows - 1
        one = np.sum(rect, 0) / counts
    out = np.tile(one, int(x.size / p) + 1)[: x.size]  # Periods.py:196-198
    if orth:
        # Periods.py:208-214: sequential, each step sees the updated projection
        for q in orth_chain(p):
            out = out - project(out, q, trunc, False)
    return out[:p] if single_period else out


def periodic_norm(x: np.ndarray, p=None) -> float:
    """RMS norm, divided by sqrt(p) for the gamma variant.  Periods.py:221-241."""
    if p:
        return (np.linalg.norm(x) / np.sqrt(len(x))) / np.sqrt(p)
    return np.linalg.norm(x) / np.sqrt(len(x))


def small_to_large(x: np.ndarray, thresh: float = 0.1, n_periods=None,
                   trunc: bool = False, orth: bool = False):
    """Periods.py:246-287.  Returns three Python lists (ragged)."""
    periods, powers, bases = [], [], []
    ref = periodic_norm(x)
    residual = x.copy()
    if n_periods is None:
        n_periods = math.floor(len(x) / 2)
    for p in range(2, n_periods + 1):
        base = project(residual, p, trunc, orth)
        trial = residual - base
        gain = (periodic_norm(residual) - periodic_norm(trial)) / ref
        if gain > thresh:
            residual = trial
            periods.append(p)
            powers.append(gain)
            bases.append(base)
    return periods, powers, bases


def fold_abs_max(x: np.ndarray, p: int) -> float:
    """max_s |sum(x[s::p])| with sequential (builtin-sum order) accumulation.

    Periods.py:327-331 uses Python's builtin `sum` over x[s::p]; an axis-0 reduce of
    the zero-padded (rows, p) rectangle adds the same terms in the same order.
    """
    rect, _ = _fold_rect(x, p)
    return float(np.max(np.abs(np.sum(rect, 0))))


def best_correlation(x: np.ndarray, num: int = 5, max_length=None, ratio: float = 0.01,
                     trunc: bool = False, orth: bool = False):
    """Periods.py:289-349."""
    if max_length is None:
        max_length = math.floor(len(x) / 3)
    periods = np.zeros(num, dtype=np.uint32)
    norms = np.zeros(num)
    bases = np.zeros((num, len(x)))
    ref = periodic_norm(x)
    prev = ref
    work = x.copy()
    for i in range(num):
        best, best_p = 0, None
        for p in range(2, max_length):  # excludes max_length, Periods.py:324
            cor = fold_abs_max(work, p)
            if cor > best:
                best, best_p = cor, p
        base = project(work, best_p, trunc, orth)  # TypeError if best_p is None, as the reference
        work = work - base
        now = periodic_norm(work)
        drop = (prev - now) / ref
        if drop > ratio:
            periods[i] = best_p
            norms[i] = drop
            bases[i] = base
            prev = now
    return periods, norms, bases


def best_frequency(x: np.ndarray, win_size=None, num: int = 5, trunc: bool = False, orth: bool = False):
    """Periods.py:351-398 (FFT peak -> period -> projection).  Listed as "next" in SURVEY.md 8f."""
    if win_size is None:
        win_size = len(x)
    periods = np.zeros(num, dtype=np.uint32)
    norms = np.zeros(num)
    bases = np.zeros((num, len(x)))
    work = x.copy()
    for i in range(num):
        mags = np.abs(np.fft.rfft(work, win_size))
        p = int(np.round((2 * win_size) / np.argmax(mags)))
        base = project(work, p, trunc, orth)
        periods[i] = p
        norms[i] = periodic_norm(base)
        bases[i] = base
        work = work - base
    return periods, norms / periodic_norm(x), bases


def m_best_meta(x: np.ndarray, gamma: bool, num: int = 5, max_length=None, min_length: int = 2,
                trunc: bool = False, orth: bool = False, stats: dict | None = None):
    """M-best / M-best-gamma.  Periods.py:456-601.

    `stats`, if given, receives {"sweeps": number of step-1 sweeps executed}.
    """
    if orth:
        warnings.warn("`Orthogonalize = True` has no effect in M-best.")  # Periods.py:482-483
    n = len(x)
    if max_length is None:
        max_length = math.floor(n / 3)
    work = x.copy()
    periods = np.zeros(num, dtype=np.uint32)
    norms = np.zeros(num)
    bases = np.zeros((num, n))
    skipped = set()

    # ---- step 1 (Periods.py:494-537)
    filled = 0
    repeats = 0
    sweeps = 0
    while filled < num:
        top, top_p, top_base = 0, 0, None
        for p in range(min_length, max_length + 1):
            base = project(work, p, trunc, orth)
            val = periodic_norm(base, p) if gamma else periodic_norm(base)
            if val > top and p not in skipped:
                top, top_p, top_base = val, p, base
        sweeps += 1
        seen = top_p in set(periods.tolist())
        if seen and repeats < 10:
            slot = np.where(periods == top_p)[0]
            bases[slot] += top_base
            norms[slot] += top
            repeats += 1
        elif seen:
            skipped.add(top_p)
            repeats = 0
        else:
            periods[filled] = top_p
            norms[filled] = top
            bases[filled] = top_base
            filled += 1
            repeats = 0
        work = work - top_base  # always, Periods.py:537 (TypeError when top_base is None)
    if stats is not None:
        stats["sweeps"] = sweeps

    # ---- step 2 (Periods.py:540-598); the outer `while changed` executes once
    stale_p = max_length  # value left in the step-1 loop variable `p` (Periods.py:559,572)
    i = 0
    while i < num:
        top, top_f, top_base, base = 0, None, None, None
        for f in factors_in_set_order(int(periods[i])):
            base = project(bases[i], f, trunc, orth)
            val = periodic_norm(base, stale_p) if gamma else periodic_norm(base)
            if val > top:
                top, top_f, top_base = val, f, base
        if top_f is not None and top_f not in periods:
            strong = top_base
            weak = bases[i] - strong
            n_strong = top
            # norm of the LAST factor's projection, not of `weak` (Periods.py:570-572)
            n_weak = periodic_norm(base, stale_p) if gamma else periodic_norm(base)
            floor = min(norms)
            if (n_weak + n_strong) > (norms[num - 1] + norms[i]) and n_weak > floor and n_strong > floor:
                bases[i] = weak
                norms[i] = n_weak
                bases = np.insert(bases, i, strong, 0)[:num]
                norms = np.insert(norms, i, n_strong)[:num]
                periods = np.insert(periods, i, top_f)[:num]
            else:
                i += 1
        else:
            i += 1

    powers = norms / periodic_norm(x)
    return periods, powers, bases


def m_best(x, num=5, max_length=None, min_length=2, trunc=False, orth=False, stats=None):
    """Periods.py:408-430."""
    return m_best_meta(x, False, num, max_length, min_length, trunc, orth, stats)


def m_best_gamma(x, num=5, max_length=None, min_length=2, trunc=False, orth=False, stats=None):
    """Periods.py:432-454."""
    return m_best_meta(x, True, num, max_length, min_length, trunc, orth, stats)
