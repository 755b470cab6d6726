"""CPU-only checks of the host side: the C-ABI library loads and exports every declared symbol,
the integer side tables agree with the oracle's, and the product never falls back to a CPU path."""
import ctypes
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib_path():
    from pyperiod_b200 import build
    return build.build()


def test_library_exports_every_declared_symbol(lib_path):
    header = open(os.path.join(ROOT, "include", "pyperiod_b200.h")).read()
    header = re.sub(r"/\*.*?\*/", "", header, flags=re.S)
    declared = set(re.findall(r"\b(pp_[a-z_0-9]+)\s*\(", header))
    assert len(declared) >= 10
    lib = ctypes.CDLL(lib_path)
    for name in sorted(declared):
        assert hasattr(lib, name), f"{name} declared in include/pyperiod_b200.h but not exported"
    from pyperiod_b200 import _lib
    assert set(_lib.SIGNATURES) == declared
    assert _lib.load().pp_abi_version() == _lib.ABI_VERSION


def test_argument_errors_do_not_need_a_gpu(lib_path):
    from pyperiod_b200 import _lib
    lib = _lib.load()
    rc = lib.pp_project(None, 8, 1, 8, 2, 0, None, 0, None, 8, 8, None)
    assert rc == -1 and b"null" in lib.pp_last_error()
    rc = lib.pp_periodic_norm(None, 1, 1, 1, 0, None, None)
    assert rc == -1


def test_tables_match_oracle_set_order():
    from oracle import numtheory
    from pyperiod_b200 import tables
    tb = tables.get_tables(1024)
    assert tb.pmax >= 1024
    for p in list(range(2, 400)) + [840, 1000, 1024]:
        assert tb.factors_of(p).tolist() == numtheory.factors_in_set_order(p)
        assert tb.chain_of(p).tolist() == numtheory.orth_chain(p) == tables.orth_chain(p)
    assert tb.factors_of(33).tolist() == [11, 3]
    assert tb.chain_of(97).tolist() == []
    big = tables.get_tables(2730)
    assert big.pmax >= 2730 and big.chain_of(2730).tolist() == numtheory.orth_chain(2730)


def test_product_does_not_import_oracle_or_fall_back():
    pkg = os.path.join(ROOT, "pyperiod_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(".py"):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), f
    from pyperiod_b200 import Periods
    import torch
    if not torch.cuda.is_available():
        with pytest.raises(RuntimeError, match="CUDA"):
            Periods.project(np.arange(10.0), 3)


def test_constructor_call_shapes():
    from pyperiod_b200 import Periods
    x = np.arange(16.0)
    a = Periods(x)                      # README form
    b = Periods(True, True)             # HEAD form
    c = Periods(x, trunc_to_integer_multiple=True)
    assert a._data is x and a.orthogonalize is False
    assert b._data is None and b.trunc_to_integer_multiple == (True, True)   # getter returns a tuple, as the reference
    assert c._trunc_to_integer_multiple is True
    b.trunc_to_integer_multiple = (False, True)
    assert b.orthogonalize is True and b._trunc_to_integer_multiple is False
    with pytest.raises(TypeError):
        Periods().m_best(num=3)         # no data anywhere


def test_synth_stream_windows_are_views():
    from pyperiod_b200 import synth
    s = synth.synth_stream(5, n=64, hop=16, segment=32)
    w = synth.windows_from_stream(s, 64, 16)
    assert w.shape == (5, 64) and w.strides == (128, 8)
    assert np.array_equal(w[3], s[48:112]) and np.max(np.abs(s)) == 1.0


def test_gcds_extracted_layout_matches_oracle_and_golden():
    """Host-built dictionary layout of QOPeriodsWithGCDsExtracted (CPython set order) == oracle == reference fixture."""
    from oracle import qo as oq
    from pyperiod_b200.qoperiods import gcds_extracted_layout
    g = np.load(os.path.join(ROOT, "tests", "golden", "qo_gcd.npz"))
    for i in range(4):
        keys, vals = g[f"c{i}_dict_keys"].tolist(), g[f"c{i}_dict_vals"].tolist()
        # the dictionary is built from ALL found periods; recover them as the keys that are not proper common factors
        found = [int(p) for p in g[f"c{i}_periods"]]
        for q in ([found] if len(found) == int(g[f"c{i}_args"][2]) else []):
            lay = gcds_extracted_layout(q)
            assert [int(k) for k in lay] == keys and list(lay.values()) == vals
    for q in ([12, 18, 30], [97], [64, 96, 40, 100], [6, 10, 15, 7]):
        lay = gcds_extracted_layout(q)
        _, ref = oq.get_subspaces_gcds_extracted(q, 200)
        assert list(lay.items()) == list(ref.items())


def _job_cover(n, pmin, pmax):
    """Independent restatement of the hierarchical job table (pp_sweep.cuh: hier_level, hier_rider_of): returns
    (jobs, how often each candidate period is evaluated)."""
    from collections import Counter

    def ctz(v):
        c = 0
        while v % 2 == 0 and c < 30:
            v //= 2
            c += 1
        return c

    lo = max(pmin, (pmax >> 1) + 1)

    def level(t):
        lv = min(ctz(t), 3)
        while lv > 0 and (t >> lv) < pmin:
            lv -= 1
        return lv

    def rider_of(q):
        lv = level(q)
        if lv < 1 or lv > 2 or ctz(q) != lv or n // (q >> lv) < (3 << lv):
            return 0
        if 3 * q <= 2 * pmax:
            r = 3 * q // 2
        elif lv == 2:
            r = 3 * q // 4
        else:
            return 0
        if r < lo or r > pmax or level(r) != ctz(r):
            return 0
        return r

    riders = {rider_of(q) for q in range(lo, pmax + 1)} - {0}
    seen, jobs = Counter(), 0
    for q in range(lo, pmax + 1):
        if q in riders:
            continue
        jobs += 1
        lv, g = level(q), q >> level(q)
        for i in range(lv + 1):
            seen[g << i] += 1
        if rider_of(q):
            for i in range(lv):
                p = (3 * g) << i
                if pmin <= p <= pmax:
                    seen[p] += 1
        if lv == 3:
            h = g
            while h % 2 == 0 and (h >> 1) >= pmin:
                h >>= 1
                seen[h] += 1
    return jobs, seen


def test_hierarchical_job_table_covers_every_candidate_once(lib_path):
    """Host mirror of the job table: pp_sweep_passes equals an independent restatement, and that restatement
    evaluates every candidate period exactly once (hosts, riders, chains)."""
    from pyperiod_b200 import _lib
    lib = _lib.load()
    assert _lib.default_fold_mode() == _lib.FOLD_HIERARCHICAL
    for n, pmin, pmax in [(4096, 2, 1024), (4096, 2, 1365), (2000, 2, 300), (4096, 2, 682), (512, 5, 64), (128, 40, 64),
                          (8192, 2, 2729), (3001, 3, 999), (2000, 17, 1000), (1000, 17, 500), (300, 2, 100),
                          (64, 2, 7), (4096, 600, 1024), (1024, 2, 512)]:
        jobs, seen = _job_cover(n, pmin, pmax)
        assert _lib.sweep_passes(n, pmin, pmax) == jobs, (n, pmin, pmax)
        assert all(seen[p] == 1 for p in range(pmin, pmax + 1)), (pmin, pmax)
        assert all(pmin <= p <= pmax for p in seen), (pmin, pmax)
    assert _lib.sweep_passes(4096, 2, 1024) == 405
    assert _lib.sweep_passes(1000, 17, 500) > _lib.sweep_passes(4096, 17, 500)   # short windows: fewer riders


def test_float_nomination_error_bound_holds():
    """The bound behind PP_FOLD_NOMINATE_F32 (pp_sweep.cuh: f32_energy_tol): |float energy - fp64 energy| of a
    candidate never exceeds 2 (2 ceil(N/p) + 40) 2^-24 sum x^2 (numpy emulation of a float fold)."""
    from pyperiod_b200 import synth
    n, u = 4096, 2.0 ** -24
    worst = 0.0
    for seed in range(2):
        x = synth.synth(n, 31_000 + seed)
        e_res = float((x * x).sum())
        xf = x.astype(np.float32)
        for p in list(range(2, 1025, 29)) + [2, 3, 5, 1023, 1024]:
            idx = np.arange(n) % p
            cnt = np.bincount(idx, minlength=p).astype(np.float64)
            s64 = np.bincount(idx, weights=x, minlength=p)
            e64 = float((s64 * s64 / cnt).sum())
            s32 = np.zeros(p, np.float32)
            m = n // p
            for k in range(m):
                s32 += xf[k * p:(k + 1) * p]
            if n - m * p:
                s32[: n - m * p] += xf[m * p:]
            e32 = float(((s32 * s32).astype(np.float64) / cnt).sum())
            tol = 2.0 * (2.0 * ((n + p - 1) // p) + 40.0) * u * e_res
            worst = max(worst, abs(e32 - e64) / tol)
    assert worst < 0.1, worst


def test_pipeline_bounds_partition_the_batch():
    """Upload pieces of the pipelined host path (pyperiod_b200/_device.py) tile [0, B) without gaps or overlap."""
    from pyperiod_b200._device import pipeline_bounds
    for b in (1, 2, 7, 16384, 16385, 131072, 1_000_003):
        bounds = pipeline_bounds(b)
        assert bounds[0][0] == 0 and bounds[-1][1] == b
        assert all(lo < hi for lo, hi in bounds)
        assert all(bounds[i][1] == bounds[i + 1][0] for i in range(len(bounds) - 1))
        assert len(bounds) <= 10


def test_threaded_host_copy_is_a_plain_copy(monkeypatch):
    """_device._host_copy (the staging-buffer -> result-array step of a pipelined download) splits a chunk over a few
    threads on 4 KB boundaries: odd lengths, lengths below the threading threshold, one thread."""
    from pyperiod_b200 import _device
    rng = np.random.default_rng(0)
    monkeypatch.setattr(_device, "COPY_THREADS_MIN_BYTES", 1 << 12)
    for n in (1, 4095, 4096, 4097, 65_537, 1_000_003):
        src = rng.integers(0, 256, n, dtype=np.uint8)
        dst = np.zeros(n, dtype=np.uint8)
        _device._host_copy(dst, src)
        assert np.array_equal(dst, src), n
    monkeypatch.setattr(_device, "COPY_THREADS", 1)
    src = rng.integers(0, 256, 70_001, dtype=np.uint8)
    dst = np.zeros_like(src)
    _device._host_copy(dst, src)
    assert np.array_equal(dst, src)
    if _device._copy_pool is not None:     # leave no worker threads behind (later tests fork)
        _device._copy_pool.shutdown()
        _device._copy_pool = None
