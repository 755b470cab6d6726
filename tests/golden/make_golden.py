#!/usr/bin/env python
"""Generate tests/golden/*.npz from the (patched) reference.  BUILD-CONTAINER ONLY.

Run:  python tests/golden/make_golden.py            (about 3-4 minutes, one core)

Every array is produced by woolgathering/pyPeriod at /root/reference loaded through
tests/golden/patched_reference.py (SURVEY.md §8c patch set), on inputs regenerated
from seeds by pyperiod_b200/synth.py.  The fixtures are what pins the oracle
(tests/test_oracle_golden.py) and, through it or directly, the CUDA path.
"""
from __future__ import annotations

import contextlib
import hashlib
import io
import os
import sys
import warnings

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

import patched_reference  # noqa: E402
from pyperiod_b200 import synth  # noqa: E402

ref = patched_reference.load()
Periods, QOPeriods, RamanujanPeriods = ref.Periods, ref.QOPeriods, ref.RamanujanPeriods


def sha(a: np.ndarray) -> str:
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def quiet(fn, *a, **k):
    with contextlib.redirect_stdout(io.StringIO()), warnings.catch_warnings():
        warnings.simplefilter("ignore")
        return fn(*a, **k)


def save(name, **arrays):
    path = os.path.join(HERE, name + ".npz")
    np.savez_compressed(path, **arrays)
    print(f"{name}.npz  {os.path.getsize(path) / 1024:.0f} KiB")


# ------------------------------------------------------------------ 1. project() vectors
def gen_project():
    out = {}
    cases = []
    for n, seed in [(200, 1), (2000, 2), (2048, 3), (4096, 4), (8192, 5), (4095, 6)]:
        x = synth.synth(n, seed)
        out[f"in_sha_{n}_{seed}"] = np.array(sha(x))
        ps = [2, 3, 7, 12, 30, 33, 64, 97, 128, 210, 255, 360, 512, 840, 1000, 1024, n // 3, n // 2]
        for p in sorted({p for p in ps if 2 <= p <= n // 2}):
            for trunc in (False, True):
                for orth in (False, True):
                    y = quiet(Periods.project, x, p, trunc, orth)
                    key = f"{n}_{seed}_{p}_{int(trunc)}{int(orth)}"
                    if orth:  # non-orth cases are pinned bit-exactly by the sha alone
                        out["one_" + key] = y[:p].copy()
                    out["sha_" + key] = np.array(sha(y))
                    out["norm_" + key] = np.array([Periods.periodic_norm(y), Periods.periodic_norm(y, p)])
                    cases.append(key)
    out["cases"] = np.array(cases)
    # tiny exact KATs
    out["kat_arange10_p3"] = Periods.project(np.arange(10.0), 3)
    out["kat_arange10_p3_trunc"] = Periods.project(np.arange(10.0), 3, True)
    out["kat_arange12_p4_orth"] = quiet(Periods.project, np.arange(12.0), 4, False, True)
    save("project", **out)


# ------------------------------------------------------------------ 2. config 1 (README signal)
def gen_readme():
    c = synth.readme_signal(0)
    out = {"in_sha": np.array(sha(c))}
    for tag, (trunc, orth) in {"00": (False, False), "11": (True, True)}.items():
        P = Periods(trunc, orth)
        per, pw, bs = quiet(P.small_to_large, c, 0.1)
        out[f"s2l_{tag}_periods"], out[f"s2l_{tag}_powers"] = np.array(per), np.array(pw)
        out[f"s2l_{tag}_bases"] = np.array(bs)
        for name, fn in (("mbest", P.m_best), ("gamma", P.m_best_gamma)):
            per, pw, bs = quiet(fn, c, num=10)
            out[f"{name}_{tag}_periods"], out[f"{name}_{tag}_powers"], out[f"{name}_{tag}_bases"] = per, pw, bs
        per, pw, bs = quiet(P.best_correlation, c, num=3)
        out[f"bcorr_{tag}_periods"], out[f"bcorr_{tag}_powers"], out[f"bcorr_{tag}_bases"] = per, pw, bs
    save("readme", **out)

    # QOPeriods / Ramanujan on the README signal
    out = {"in_sha": np.array(sha(c))}
    q = QOPeriods()
    d, res = quiet(q.find_periods, c, num=2, thresh=0.05)
    out.update(qo_periods=np.array(d["periods"]), qo_norms=np.array(d["norms"]), qo_weights=d["weights"],
               qo_res=res, qo_dict_keys=np.array([int(k) for k in d["basis_dictionary"]]),
               qo_dict_vals=np.array(list(d["basis_dictionary"].values())))
    gp = quiet(q.get_periods, d["weights"], d["basis_dictionary"], "lstsq")
    for i, g in enumerate(gp):
        out[f"qo_getp_lstsq_{i}"] = g
    d3, res3 = quiet(q.find_periods, c, num=3, thresh=0.05)
    out.update(qo3_periods=np.array(d3["periods"]), qo3_norms=np.array(d3["norms"]), qo3_weights=d3["weights"],
               qo3_res=res3, qo3_dict_keys=np.array([int(k) for k in d3["basis_dictionary"]]),
               qo3_dict_vals=np.array(list(d3["basis_dictionary"].values())))
    for kind in ("row reduction", "lstsq"):
        try:
            gp = quiet(q.get_periods, d3["weights"], d3["basis_dictionary"], kind)
            for i, g in enumerate(gp):
                out[f"qo3_getp_{kind.replace(' ', '')}_{i}"] = g
        except np.linalg.LinAlgError:
            out[f"qo3_getp_{kind.replace(' ', '')}_linalgerror"] = np.array(1)
    r = RamanujanPeriods()
    norms = quiet(r.find_periods, c, 2, 120)
    d, res = quiet(r.find_periods_with_weights, c, max_length=120, thresh=0.2)
    out.update(ram_norms=norms, ram_periods=np.array(d["periods"]), ram_sel_norms=np.array(d["norms"]),
               ram_weights=d["weights"], ram_res=res,
               ram_dict_keys=np.array([int(k) for k in d["basis_dictionary"]]),
               ram_dict_vals=np.array(list(d["basis_dictionary"].values())))
    out["cq6"] = RamanujanPeriods.Cq(6)
    out["cq12"] = RamanujanPeriods.Cq(12)
    out["cq30"] = RamanujanPeriods.Cq(30)
    save("readme_qo_ram", **out)


# ------------------------------------------------------------------ 3. config 3: M-best on stream windows
def gen_mbest_stream():
    stream = synth.synth_stream(n_windows=128 * 3 + 1, n=4096, hop=512, seed0=30_000)
    win = synth.windows_from_stream(stream)
    out = {"stream_sha": np.array(sha(stream)), "window_ids": np.array([0, 128, 300])}
    P = Periods()
    for b in (0, 128, 300):
        x = np.array(win[b])
        for name, fn in (("mbest", P.m_best), ("gamma", P.m_best_gamma)):
            per, pw, bs = quiet(fn, x, num=10, max_length=1024)
            out[f"{name}_{b}_periods"], out[f"{name}_{b}_powers"] = per, pw
            out[f"{name}_{b}_bases_sha"] = np.array(sha(bs))
            out[f"{name}_{b}_bases_one"] = np.concatenate([bs[i, : int(per[i])] for i in range(10)])
    save("mbest_stream", **out)


# ------------------------------------------------------------------ 4. config 2: small-to-large, N=2048
def gen_s2l():
    out = {}
    for b in range(4):
        x = synth.synth(2048, 20_000 + b)
        for tag, (trunc, orth) in {"00": (False, False), "10": (True, False), "11": (True, True)}.items():
            per, pw, bs = quiet(Periods(trunc, orth).small_to_large, x, 0.1)
            out[f"s2l_{b}_{tag}_periods"], out[f"s2l_{b}_{tag}_powers"] = np.array(per), np.array(pw)
            out[f"s2l_{b}_{tag}_bases_sha"] = np.array(sha(np.array(bs)))
    save("s2l_2048", **out)


# ------------------------------------------------------------------ 5. config 4: best_correlation
def gen_bcorr():
    out = {}
    for b in range(3):
        x = synth.synth(2048, 40_000 + b)
        for tag, (trunc, orth) in {"00": (False, False), "11": (True, True)}.items():
            per, pw, bs = quiet(Periods(trunc, orth).best_correlation, x, num=5)
            out[f"n2048_{b}_{tag}_periods"], out[f"n2048_{b}_{tag}_powers"] = per, pw
            out[f"n2048_{b}_{tag}_bases_sha"] = np.array(sha(bs))
    x = synth.synth(8192, 40_000)
    per, pw, bs = quiet(Periods(True, True).best_correlation, x, num=2)
    out["n8192_0_11_periods"], out["n8192_0_11_powers"] = per, pw
    out["n8192_0_11_bases_sha"] = np.array(sha(bs))
    out["n8192_0_11_bases_one"] = np.concatenate([bs[i, : int(per[i])] for i in range(2)])
    save("bcorr", **out)


# ------------------------------------------------------------------ 6. config 5: QO + Ramanujan on synth windows
def gen_qo_ram():
    out = {}
    for b in range(2):
        x = synth.synth(4096, 50_000 + b)
        d, res = quiet(QOPeriods().find_periods, x, num=4, thresh=0.05)
        out[f"qo_{b}_periods"], out[f"qo_{b}_norms"] = np.array(d["periods"]), np.array(d["norms"])
        out[f"qo_{b}_weights"], out[f"qo_{b}_res"] = d["weights"], res
        out[f"qo_{b}_dict_keys"] = np.array([int(k) for k in d["basis_dictionary"]])
        out[f"qo_{b}_dict_vals"] = np.array(list(d["basis_dictionary"].values()))
    for b in range(2):
        x = synth.synth(1024, 50_000 + b)
        r = RamanujanPeriods()
        norms = quiet(r.find_periods, x)  # q = 2..341
        d, res = quiet(r.find_periods_with_weights, x, thresh=0.2)
        out[f"ram_{b}_norms"] = norms
        out[f"ram_{b}_periods"], out[f"ram_{b}_sel_norms"] = np.array(d["periods"]), np.array(d["norms"])
        out[f"ram_{b}_weights"], out[f"ram_{b}_res"] = d["weights"], res
        out[f"ram_{b}_dict_keys"] = np.array([int(k) for k in d["basis_dictionary"]])
        out[f"ram_{b}_dict_vals"] = np.array(list(d["basis_dictionary"].values()))
    save("qo_ram_synth", **out)


# ------------------------------------------------------------------ 7. QOPeriodsWithGCDsExtracted
def gen_qo_gcd():
    import importlib
    gcd_cls = importlib.import_module(ref.__name__ + ".QOPeriodsWithGCDsExtracted").QOPeriodsWithGCDsExtracted
    out = {}
    cases = [(8800, 1500, dict(num=3, thresh=0.05, max_length=300)), (8801, 1500, dict(num=4, thresh=0.05, max_length=300)),
             (8802, 1200, dict(num=4, thresh=0.02, max_length=200)), (50_000, 2048, dict(num=4, thresh=0.05))]
    for i, (seed, n, kw) in enumerate(cases):
        x = synth.synth(n, seed)
        d, res = quiet(gcd_cls().find_periods, x, **kw)
        out[f"c{i}_args"] = np.array([seed, n, kw["num"], kw.get("max_length", 0)])
        out[f"c{i}_thresh"] = np.array(kw["thresh"])
        out[f"c{i}_periods"], out[f"c{i}_norms"] = np.array(d["periods"]), np.array(d["norms"])
        out[f"c{i}_weights"], out[f"c{i}_res"] = d["weights"], res
        out[f"c{i}_dict_keys"] = np.array([int(k) for k in d["basis_dictionary"]])
        out[f"c{i}_dict_vals"] = np.array([int(v) for v in d["basis_dictionary"].values()])
        out[f"c{i}_subspaces_sha"] = np.array(sha(np.asarray(d["subspaces"], dtype=np.float64)))
    save("qo_gcd", **out)


# ------------------------------------------------------------------ 8. config 5 at its own shape (N = 4096, q <= 1365)
CFG5_WINDOWS = (6, 14, 16)   # dictionaries of 1234, 2896 and 1324 rows (more than the 1024 the first launch holds)


def _ram_cfg5_one(b):
    x = synth.synth(4096, 50_000 + b)
    r = RamanujanPeriods()
    d, res = quiet(r.find_periods_with_weights, x, thresh=0.2)   # q = 2..1365; about 25 minutes per window
    return b, d, res


def gen_ram_cfg5():
    """RamanujanPeriods.find_periods_with_weights(thresh=0.2) on three config-5 windows whose dictionaries have more
    than 1024 rows (about 25 minutes per window for the reference; the windows run in parallel processes)."""
    import multiprocessing as mp
    out = {"windows": np.array(CFG5_WINDOWS)}
    with mp.get_context("fork").Pool(len(CFG5_WINDOWS)) as pool:
        for b, d, res in pool.map(_ram_cfg5_one, CFG5_WINDOWS):
            out[f"w{b}_periods"], out[f"w{b}_sel_norms"] = np.array(d["periods"]), np.array(d["norms"])
            out[f"w{b}_weights"], out[f"w{b}_res"] = d["weights"], res
            out[f"w{b}_dict_keys"] = np.array([int(k) for k in d["basis_dictionary"]])
            out[f"w{b}_dict_vals"] = np.array(list(d["basis_dictionary"].values()))
    save("ram_cfg5", **out)


# ------------------------------------------------------------------ 9. tie-heavy inputs (exactly periodic, integer valued)
def tie_inputs(n=4096):
    """Inputs on which p, 2p, 3p ... give bit-identical projections, so the reference's norms tie EXACTLY and its
    strict '>' keeps the lowest period (Periods.py:512-515).  Regenerated from seeds by the tests."""
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


def gen_ties():
    out = {}
    for name, x in tie_inputs().items():
        if not x.any():
            continue
        for tag, fn in (("norm", Periods().m_best), ("gamma", Periods().m_best_gamma)):
            per, pw, bs = quiet(fn, x, num=1, max_length=1024)
            out[f"{name}_{tag}_period"], out[f"{name}_{tag}_power"] = np.array(per), np.array(pw)
            out[f"{name}_{tag}_base_sha"] = np.array(sha(bs))
    save("ties", **out)


# ------------------------------------------------------------------ 10. QOPeriods(basis_type="ramanujan") + get_periods
QO_RAMBASIS_CASES = [(600, 5, 2, None), (600, 5, 3, None), (1024, 50_001, 3, None), (2000, 7, 4, 300)]
GETP_CASES = [(2000, 7, 4, 300), (1024, 50_001, 3, 200), (4096, 50_003, 4, None)]


def gen_qo_rambasis():
    """QOPeriods(basis_type="ramanujan").find_periods (QOPeriods.py:970-971, 1005-1052): periods, dictionary, norms
    and residual (the weights solve a singular system and are not reproducible); and QOPeriods.get_periods in its
    "row reduction" and lstsq branches on natural-basis results with 3-4 periods."""
    out = {}
    for i, (n, seed, num, ml) in enumerate(QO_RAMBASIS_CASES):
        x = synth.synth(n, seed)
        q = QOPeriods(basis_type="ramanujan")
        d, res = quiet(q.find_periods, x, num=num, thresh=0.05, max_length=ml)
        out[f"c{i}_periods"], out[f"c{i}_norms"], out[f"c{i}_res"] = np.array(d["periods"]), np.array(d["norms"]), res
        out[f"c{i}_dict_keys"] = np.array([int(k) for k in d["basis_dictionary"]])
        out[f"c{i}_dict_vals"] = np.array([int(v) for v in d["basis_dictionary"].values()])
        out[f"c{i}_weights"] = d["weights"]
    for i, (n, seed, num, ml) in enumerate(GETP_CASES):
        x = synth.synth(n, seed)
        q = QOPeriods()
        d, res = quiet(q.find_periods, x, num=num, thresh=0.01, max_length=ml)
        out[f"g{i}_dict_keys"] = np.array([int(k) for k in d["basis_dictionary"]])
        out[f"g{i}_dict_vals"] = np.array([int(v) for v in d["basis_dictionary"].values()])
        out[f"g{i}_weights"] = d["weights"]
        for kind in ("row reduction", "lstsq"):
            try:
                gp = quiet(q.get_periods, d["weights"], d["basis_dictionary"], kind)
                out[f"g{i}_{kind.replace(' ', '')}"] = np.concatenate(gp)
            except np.linalg.LinAlgError:
                out[f"g{i}_{kind.replace(' ', '')}_linalgerror"] = np.array(1)
    save("qo_rambasis", **out)


# ------------------------------------------------------------------ 11. Muresan eq. 3 finder (QOPeriods.py:1122-1232)
MURESAN_CASES = [(600, 5, None), (1024, 50_001, 300), (2000, 7, None)]


def gen_muresan():
    out = {}
    q = QOPeriods()
    for i, (n, seed, max_p) in enumerate(MURESAN_CASES):
        x = synth.synth(n, seed)
        for norm in (False, True):
            out[f"c{i}_pows_{int(norm)}"] = quiet(q.get_best_period_orthogonal, x, max_p, norm, True)
            out[f"c{i}_best_{int(norm)}"] = np.array(quiet(q.get_best_period_orthogonal, x, max_p, norm, False))
        out[f"c{i}_eq3"] = np.array([quiet(q.eq_3, x, p) for p in (1, 2, 7, 30, 97, n // 4)])
    save("muresan", **out)


if __name__ == "__main__":
    which = sys.argv[1:] or ["project", "readme", "mbest_stream", "s2l", "bcorr", "qo_ram", "qo_gcd", "qo_rambasis"]   # + "ram_cfg5" (slow)
    for w in which:
        globals()["gen_" + w]()
