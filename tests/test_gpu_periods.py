"""GPU parity tests for the `Periods` path: CUDA (through the C ABI) vs the oracle and vs the
golden fixtures generated from the reference.  Run with `-m gpu` on a B200.

Bars: project() bit-exact; period lists exact; powers/bases <= 1e-10 relative (they are in fact
bit-exact for bases in the non-trivial cases below, which is asserted where it holds by design).
"""
import warnings

import numpy as np
import pytest

from conftest import load_golden, sha
from oracle import periods as op
from pyperiod_b200 import synth

pytestmark = pytest.mark.gpu

RTOL = 1e-10  # north-star tolerance for powers and bases (fp64 path)


@pytest.fixture(scope="module")
def P():
    from pyperiod_b200 import Periods
    return Periods


def _native_loaded():
    with open("/proc/self/maps") as fh:
        return "libpyperiod_b200.so" in fh.read()


def test_native_library_is_what_runs(P):
    P.project(synth.synth(256, 1), 7)
    assert _native_loaded()


# ---------------------------------------------------------------- project: bit-exact
def test_project_kats(P):
    assert np.array_equal(P.project(np.arange(10.0), 3), [4.5, 4, 5, 4.5, 4, 5, 4.5, 4, 5, 4.5])
    assert np.array_equal(P.project(np.arange(10.0), 3, True), [3, 4, 5, 3, 4, 5, 3, 4, 5, 3])
    assert np.array_equal(P.project(np.arange(12.0), 4, False, True), [-1, -1, 1, 1] * 3)
    assert np.array_equal(P.project(np.arange(12.0), 4, False, True, True), [-1, -1, 1, 1])


def test_project_golden_bit_exact(P):
    g = load_golden("project")
    cache = {}
    for key in g["cases"]:
        n, seed, p, mode = str(key).split("_")
        n, seed, p = int(n), int(seed), int(p)
        x = cache.setdefault((n, seed), synth.synth(n, seed))
        y = P.project(x, p, mode[0] == "1", mode[1] == "1")
        assert sha(y) == str(g["sha_" + str(key)]), key


def test_project_all_periods_vs_oracle(P):
    x = synth.synth(1000, 77)
    for p in range(2, 501):
        for trunc, orth in ((False, False), (True, False), (False, True), (True, True)):
            got = P.project(x, p, trunc, orth)
            assert np.array_equal(got, op.project(x, p, trunc, orth)), (p, trunc, orth)


def test_project_batch_and_ragged_lengths(P):
    for n in (257, 1000, 4095, 4096):
        xb = synth.synth_batch(5, n, 900)
        for p, trunc, orth in ((3, False, False), (30, True, True), (n // 2, False, True), (n, False, False)):
            got = P.project(xb, p, trunc, orth)
            for b in range(5):
                assert np.array_equal(got[b], op.project(xb[b], p, trunc, orth)), (n, p, b)
    one = P.project(synth.synth(300, 5), 12, True, True, True)
    assert one.shape == (12,)


def test_project_idempotent_and_periodic(P):
    x = synth.synth(4096, 11)
    for p in (5, 64, 1000):
        y = P.project(x, p)
        assert np.array_equal(y[: 4096 - p], y[p:])           # p-periodic
        assert np.allclose(P.project(y, p), y, rtol=0, atol=1e-15)  # idempotent


def test_periodic_norm(P):
    x = synth.synth(2000, 3)
    assert abs(P.periodic_norm(x) - op.periodic_norm(x)) <= 1e-15
    assert abs(P.periodic_norm(x, 7) - op.periodic_norm(x, 7)) <= 1e-15


# ---------------------------------------------------------------- sweep: energies and argmax
@pytest.mark.parametrize("trunc,orth", [(False, False), (True, False), (False, True), (True, True)])
def test_sweep_metrics_vs_oracle(P, trunc, orth):
    x = synth.synth(2000, 21)
    pmax = 666
    for metric in ("norm", "gamma"):
        m, bp, bv = P(trunc, orth).sweep(x[None, :], metric=metric, max_length=pmax)
        want = np.zeros(pmax + 1)
        for p in range(2, pmax + 1):
            base = op.project(x, p, trunc, orth)
            want[p] = op.periodic_norm(base, p) if metric == "gamma" else op.periodic_norm(base)
        np.testing.assert_allclose(m[0, 2:], want[2:], rtol=1e-12, atol=1e-15)
        assert int(bp[0]) == int(np.argmax(want))
    # imposed norm (small-to-large metric)
    m, _, _ = P(trunc, orth).sweep(x[None, :], metric="imposed", max_length=pmax)
    ref = op.periodic_norm(x)
    for p in (2, 3, 58, 59, 100, 333, 666):
        base = op.project(x, p, trunc, orth)
        want_p = (op.periodic_norm(x) - op.periodic_norm(x - base)) / ref
        assert abs(m[0, p] - want_p) <= 1e-13, p


@pytest.mark.parametrize("n,pmin,pmax", [(4096, 2, 1024), (2000, 2, 666), (4095, 3, 1000), (1000, 17, 500),
                                          (8192, 2, 2730), (512, 2, 100), (4096, 600, 1024), (4096, 2, 1365),
                                          (3001, 2, 1500), (777, 5, 259), (4096, 2, 682), (1024, 2, 341)])
def test_sweep_hierarchical_matches_direct(P, n, pmin, pmax):
    """The hierarchical ranking sweep (S_p from S_2p, tops that are 3/2 or 3/4 of an even top riding on its
    pass) agrees with the sequential fold to rounding, for every candidate period."""
    from pyperiod_b200 import _lib
    xb = synth.synth_batch(3, n, 4242)
    try:
        out = {}
        for mode in (_lib.FOLD_DIRECT, _lib.FOLD_HIERARCHICAL_NO_RIDERS, _lib.FOLD_HIERARCHICAL):
            _lib.set_fold_mode(mode)
            for metric in ("norm", "gamma"):
                out[mode, metric] = P().sweep(xb, metric=metric, min_length=pmin, max_length=pmax)
    finally:
        _lib.set_fold_mode(_lib.FOLD_HIERARCHICAL)
    for metric in ("norm", "gamma"):
        md, pd_, vd = out[_lib.FOLD_DIRECT, metric]
        mh, ph, vh = out[_lib.FOLD_HIERARCHICAL, metric]
        assert np.all(mh[:, :pmin] == 0) and np.all(md[:, :pmin] == 0)
        np.testing.assert_allclose(mh[:, pmin:], md[:, pmin:], rtol=1e-13, atol=0)
        assert np.array_equal(pd_, ph)
        np.testing.assert_allclose(vh, vd, rtol=1e-13)
        mn, pn, vn = out[_lib.FOLD_HIERARCHICAL_NO_RIDERS, metric]
        np.testing.assert_allclose(mn[:, pmin:], md[:, pmin:], rtol=1e-13, atol=0)
        assert np.array_equal(pd_, pn)


@pytest.mark.parametrize("n,pmin,pmax", [(4096, 2, 1024), (2000, 2, 666), (4095, 3, 1000), (1000, 17, 500),
                                          (8192, 2, 2730), (512, 2, 100), (4096, 600, 1024), (4096, 2, 1365),
                                          (3001, 2, 1500), (777, 5, 259), (4099, 2, 2049), (1024, 2, 341)])
def test_sweep_hierarchical_truncated_matches_direct(P, n, pmin, pmax):
    """trunc_to_integer_multiple=True: the hierarchical sweep (accumulator sets corrected by the rows a level does
    not use, an extra complete row where the halved period has an odd row count) agrees with the sequential
    truncated fold to rounding for every candidate period, and picks the same period."""
    from pyperiod_b200 import _lib
    xb = synth.synth_batch(3, n, 4343)
    try:
        out = {}
        for mode in (_lib.FOLD_DIRECT, _lib.FOLD_HIERARCHICAL):
            _lib.set_fold_mode(mode)
            for metric in ("norm", "gamma"):
                out[mode, metric] = P(True, False).sweep(xb, metric=metric, min_length=pmin, max_length=pmax)
    finally:
        _lib.set_fold_mode(_lib.FOLD_HIERARCHICAL)
    for metric in ("norm", "gamma"):
        md, pd_, vd = out[_lib.FOLD_DIRECT, metric]
        mh, ph, vh = out[_lib.FOLD_HIERARCHICAL, metric]
        assert np.all(mh[:, :pmin] == 0)
        np.testing.assert_allclose(mh[:, pmin:], md[:, pmin:], rtol=1e-12, atol=0)
        assert np.array_equal(pd_, ph)
        np.testing.assert_allclose(vh, vd, rtol=1e-12)
    # and against the oracle's projection for one window
    x = xb[0]
    for p in sorted({pmin, pmin + 1, (pmin + pmax) // 2, pmax - 1, pmax}):
        want = op.periodic_norm(op.project(x, p, True, False))
        assert abs(out[_lib.FOLD_HIERARCHICAL, "norm"][0][0, p] - want) <= 1e-12 * max(want, 1e-300), p


def test_mbest_truncated_fold_modes_agree(P):
    """Periods(True, False).m_best ranks hierarchically: same periods, sweep counts and bit-identical bases as with
    sequential folds."""
    from pyperiod_b200 import _lib
    xb = synth.synth_batch(8, 4096, 778)
    for gamma in (False, True):
        try:
            _lib.set_fold_mode(_lib.FOLD_DIRECT)
            fn = P(True, False).m_best_gamma if gamma else P(True, False).m_best
            a = fn(xb, num=10, max_length=1024, return_bases=True)
            _lib.set_fold_mode(_lib.FOLD_HIERARCHICAL)
            b = fn(xb, num=10, max_length=1024, return_bases=True)
        finally:
            _lib.set_fold_mode(_lib.FOLD_HIERARCHICAL)
        assert np.array_equal(a.periods, b.periods) and np.array_equal(a.sweeps, b.sweeps)
        np.testing.assert_allclose(a.powers, b.powers, rtol=1e-12)
        assert np.array_equal(a.bases, b.bases)


def test_mbest_fold_modes_agree(P):
    from pyperiod_b200 import _lib
    xb = synth.synth_batch(8, 4096, 777)
    try:
        _lib.set_fold_mode(_lib.FOLD_DIRECT)
        a = P().m_best_gamma(xb, num=10, max_length=1024, return_bases=True)
        _lib.set_fold_mode(_lib.FOLD_HIERARCHICAL)
        b = P().m_best_gamma(xb, num=10, max_length=1024, return_bases=True)
    finally:
        _lib.set_fold_mode(_lib.FOLD_HIERARCHICAL)
    assert np.array_equal(a.periods, b.periods) and np.array_equal(a.sweeps, b.sweeps)
    np.testing.assert_allclose(a.powers, b.powers, rtol=1e-12)
    assert np.array_equal(a.bases, b.bases)       # bases never come from the hierarchical sums


def test_sweep_maxabs_bit_exact(P):
    x = synth.synth(2048, 40_000)
    m, bp, bv = P().sweep(x[None, :], metric="maxabs", max_length=681)
    want = np.array([0, 0] + [op.fold_abs_max(x, p) for p in range(2, 682)])
    assert np.array_equal(m[0, 2:], want[2:])
    assert int(bp[0]) == int(np.argmax(want)) and bv[0] == want.max()


# ---------------------------------------------------------------- config 1: README signal
@pytest.mark.parametrize("tag", ["00", "11"])
def test_readme_algorithms_vs_golden(P, tag):
    g = load_golden("readme")
    c = synth.readme_signal(0)
    trunc, orth = tag[0] == "1", tag[1] == "1"
    inst = P(c, trunc, orth)  # README call shape
    per, pw, bs = inst.small_to_large(thresh=0.1)
    assert per == g[f"s2l_{tag}_periods"].tolist()
    np.testing.assert_allclose(pw, g[f"s2l_{tag}_powers"], rtol=RTOL)
    np.testing.assert_allclose(np.array(bs), g[f"s2l_{tag}_bases"], rtol=RTOL, atol=1e-14)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        for name, fn in (("mbest", inst.m_best), ("gamma", inst.m_best_gamma)):
            per, pw, bs = fn(num=10)
            assert per.dtype == np.uint32 and np.array_equal(per, g[f"{name}_{tag}_periods"]), name
            np.testing.assert_allclose(pw, g[f"{name}_{tag}_powers"], rtol=RTOL)
            np.testing.assert_allclose(bs, g[f"{name}_{tag}_bases"], rtol=RTOL, atol=1e-14)
    per, pw, bs = inst.best_correlation(num=3)
    assert np.array_equal(per, g[f"bcorr_{tag}_periods"])
    np.testing.assert_allclose(pw, g[f"bcorr_{tag}_powers"], rtol=RTOL, atol=1e-15)
    np.testing.assert_allclose(bs, g[f"bcorr_{tag}_bases"], rtol=RTOL, atol=1e-14)


def test_readme_bases_bit_exact_nonorth(P):
    g = load_golden("readme")
    c = synth.readme_signal(0)
    per, pw, bs = P().m_best(c, num=10)  # HEAD call shape
    assert np.array_equal(bs, g["mbest_00_bases"])
    per, pw, bs = P().best_correlation(c, num=3)
    assert np.array_equal(bs, g["bcorr_00_bases"])
    per, pw, bs = P().small_to_large(c, 0.1)
    assert np.array_equal(np.array(bs), g["s2l_00_bases"])


# ---------------------------------------------------------------- config 3: stream windows, N=4096
def test_mbest_stream_windows_vs_golden(P):
    g = load_golden("mbest_stream")
    stream = synth.synth_stream(n_windows=128 * 3 + 1)
    win = synth.windows_from_stream(stream)        # overlapping strided view, uploaded once
    inst = P()
    for name, fn in (("mbest", inst.m_best), ("gamma", inst.m_best_gamma)):
        res = fn(win, num=10, max_length=1024, return_bases=True)
        assert res.status.max() == 0
        for b in (0, 128, 300):
            assert np.array_equal(res.periods[b], g[f"{name}_{b}_periods"]), (name, b)
            np.testing.assert_allclose(res.powers[b], g[f"{name}_{b}_powers"], rtol=RTOL)
            assert sha(res.bases[b]) == str(g[f"{name}_{b}_bases_sha"]), (name, b)
        assert res.sweeps.min() >= 10


def test_mbest_batch_vs_oracle_and_loop(P):
    xb = synth.synth_batch(6, 2048, 5000)
    inst = P()
    for gamma in (False, True):
        fn = inst.m_best_gamma if gamma else inst.m_best
        res = fn(xb, num=8, max_length=600, return_bases=True)
        for b in range(6):
            st = {}
            per, pw, bs = op.m_best_meta(xb[b], gamma, 8, 600, 2, stats=st)
            assert np.array_equal(res.periods[b], per), (gamma, b)
            np.testing.assert_allclose(res.powers[b], pw, rtol=RTOL)
            np.testing.assert_allclose(res.bases[b], bs, rtol=RTOL, atol=1e-14)
            assert int(res.sweeps[b]) == st["sweeps"]
            one = fn(xb[b], num=8, max_length=600)               # (N,) call == row of the batch
            assert np.array_equal(one[0], res.periods[b]) and np.array_equal(one[1], res.powers[b])
            assert np.array_equal(one[2], res.bases[b])


def test_mbest_trunc_modes_vs_oracle(P):
    x = synth.synth(2000, 6001)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        for trunc, orth in ((True, False), (True, True), (False, True)):
            for gamma in (False, True):
                fn = P(trunc, orth).m_best_gamma if gamma else P(trunc, orth).m_best
                per, pw, bs = fn(x, num=6, max_length=400)
                per0, pw0, bs0 = op.m_best_meta(x, gamma, 6, 400, 2, trunc, orth)
                assert np.array_equal(per, per0), (trunc, orth, gamma)
                np.testing.assert_allclose(pw, pw0, rtol=RTOL)
                np.testing.assert_allclose(bs, bs0, rtol=RTOL, atol=1e-14)


def test_mbest_zero_input_raises(P):
    with pytest.raises(TypeError):
        P().m_best(np.zeros(512), num=3)
    res = P().m_best(np.zeros((2, 512)), num=3)
    assert res.status.tolist() == [1, 1]


# ---------------------------------------------------------------- config 2: small-to-large
def test_s2l_vs_golden(P):
    g = load_golden("s2l_2048")
    xb = synth.synth_batch(4, 2048, 20_000)
    for tag in ("00", "10", "11"):
        res = P(tag[0] == "1", tag[1] == "1").small_to_large(xb, thresh=0.1, return_bases=True)
        for b in range(4):
            per, pw, bs = res.window(b)
            assert per == g[f"s2l_{b}_{tag}_periods"].tolist(), (tag, b)
            np.testing.assert_allclose(pw, g[f"s2l_{b}_{tag}_powers"], rtol=RTOL)
            if tag == "00":
                assert sha(np.array(bs)) == str(g[f"s2l_{b}_{tag}_bases_sha"])


def test_s2l_kmax_overflow_and_rerun(P):
    x = synth.synth(1024, 31)
    per, pw, bs = P().small_to_large(x, thresh=0.001, kmax=2)   # 1-D: reruns with enough room
    per0, pw0, _ = op.small_to_large(x, 0.001)
    assert per == per0 and len(per) > 2
    res = P().small_to_large(x[None, :].repeat(2, 0), thresh=0.001, kmax=2)
    assert res.status.tolist() == [2, 2] and res.count.tolist() == [len(per0)] * 2
    assert res.periods[0].tolist() == per0[:2]


# ---------------------------------------------------------------- config 4: best correlation
def test_bcorr_vs_golden(P):
    g = load_golden("bcorr")
    xb = synth.synth_batch(3, 2048, 40_000)
    for tag in ("00", "11"):
        res = P(tag[0] == "1", tag[1] == "1").best_correlation(xb, num=5, return_bases=True)
        for b in range(3):
            assert np.array_equal(res.periods[b], g[f"n2048_{b}_{tag}_periods"]), (tag, b)
            np.testing.assert_allclose(res.powers[b], g[f"n2048_{b}_{tag}_powers"], rtol=RTOL, atol=1e-15)
            assert sha(res.bases[b]) == str(g[f"n2048_{b}_{tag}_bases_sha"]), (tag, b)
    x = synth.synth(8192, 40_000)
    per, pw, bs = P(True, True).best_correlation(x, num=2)
    assert np.array_equal(per, g["n8192_0_11_periods"])
    np.testing.assert_allclose(pw, g["n8192_0_11_powers"], rtol=RTOL)
    assert sha(bs) == str(g["n8192_0_11_bases_sha"])


# ---------------------------------------------------------------- device-resident input / output
def test_device_tensor_io(P):
    import torch
    x = torch.from_numpy(synth.synth_batch(3, 1024, 7)).cuda()
    res = P().m_best(x, num=4, max_length=300)
    assert res.periods.is_cuda and res.bases is None
    host = P().m_best(x.cpu().numpy(), num=4, max_length=300)
    assert np.array_equal(res.periods.cpu().numpy().view(np.uint32), host.periods)
    assert np.array_equal(res.powers.cpu().numpy(), host.powers)


# ---------------------------------------------------------------- edge cases
def test_empty_batch_and_minimal_sizes(P):
    empty = np.zeros((0, 512))
    res = P().m_best(empty, num=3, max_length=100)
    assert res.periods.shape == (0, 3) and res.powers.shape == (0, 3)
    r2 = P().small_to_large(empty, thresh=0.1)
    assert r2.count.shape == (0,)
    assert P.project(empty, 5).shape == (0, 512)
    # smallest useful window, period equal to the window length, odd length (no TMA alignment)
    x = synth.synth(7, 1)
    for p in (1, 2, 3, 7):
        got = P.project(x, p)
        if p > 1:   # p = 1 is numpy's pairwise-summation case
            assert np.array_equal(got, op.project(x, p)), p
        else:
            np.testing.assert_allclose(got, op.project(x, p), rtol=1e-15)
    # p > N: the reference pads to one row -- the projection is the data itself, nan with truncation
    assert np.array_equal(P.project(x, 8), x) and np.array_equal(P.project(x, 8, return_single_period=True), x)
    assert np.isnan(P.project(x, 8, True)).all()
    with pytest.raises(ValueError):
        P.project(x, 0)


def test_scaled_and_offset_windows(P):
    """Amplitude and DC offset do not change which periods are found (ranking is scale-free), and the
    batch path treats every row independently."""
    x = synth.synth(2048, 77)
    xb = np.stack([x, 1e6 * x, 1e-6 * x, x + 3.0])
    res = P().m_best_gamma(xb, num=5, max_length=500)
    for b in range(4):
        per, pw, _ = op.m_best_gamma(xb[b], 5, 500)
        assert np.array_equal(res.periods[b], per), b
        np.testing.assert_allclose(res.powers[b], pw, rtol=RTOL)
    assert np.array_equal(res.periods[0], res.periods[1]) and np.array_equal(res.periods[0], res.periods[2])


def test_large_window_n8192_mbest(P):
    x = synth.synth(8192, 40_001)
    per, pw, bs = P().m_best(x, num=4)                # default max_length = 2730: multi-tile tops
    per0, pw0, bs0 = op.m_best(x, 4)
    assert np.array_equal(per, per0)
    np.testing.assert_allclose(pw, pw0, rtol=RTOL)
    assert np.array_equal(bs, bs0)


def test_best_frequency_vs_oracle(P):
    """Periods.best_frequency (SURVEY.md 8f 'next' #1): cuFFT spectrum + this library's exact projection."""
    for seed, (trunc, orth) in ((901, (False, False)), (902, (True, True))):
        x = synth.synth(2000, seed)
        per, pw, bs = P(trunc, orth).best_frequency(x, num=4)
        per0, pw0, bs0 = op.best_frequency(x, None, 4, trunc, orth)
        assert np.array_equal(per, per0)
        np.testing.assert_allclose(pw, pw0, rtol=RTOL)
        np.testing.assert_allclose(bs, bs0, rtol=RTOL, atol=1e-14)
    xb = synth.synth_batch(3, 1024, 950)
    res = P().best_frequency(xb, num=3)
    for b in range(3):
        per0, pw0, _ = op.best_frequency(xb[b], None, 3)
        assert np.array_equal(res.periods[b], per0)
        np.testing.assert_allclose(res.powers[b], pw0, rtol=RTOL)


# ---------------------------------------------------------------- randomized differential sweep
@pytest.mark.parametrize("trunc", [False, True])
def test_differential_many_windows(P, trunc):
    """48 random windows x 4 algorithms against the oracle: period lists exact, powers <= 1e-10."""
    xb = synth.synth_batch(48, 1024, 777_000 + int(trunc))
    inst = P(trunc, False)
    mb = inst.m_best(xb, num=6, max_length=300)
    mg = inst.m_best_gamma(xb, num=6, max_length=300)
    sl = inst.small_to_large(xb, thresh=0.08)
    bc = inst.best_correlation(xb, num=4, max_length=300)
    for b in range(48):
        per, pw, _ = op.m_best(xb[b], 6, 300, 2, trunc)
        assert np.array_equal(mb.periods[b], per), ("m_best", b)
        np.testing.assert_allclose(mb.powers[b], pw, rtol=RTOL)
        per, pw, _ = op.m_best_gamma(xb[b], 6, 300, 2, trunc)
        assert np.array_equal(mg.periods[b], per), ("gamma", b)
        np.testing.assert_allclose(mg.powers[b], pw, rtol=RTOL)
        per, pw, _ = op.small_to_large(xb[b], 0.08, None, trunc)
        k = int(sl.count[b])
        assert sl.periods[b, :k].tolist() == per, ("s2l", b)
        np.testing.assert_allclose(sl.powers[b, :k], pw, rtol=1e-9)
        per, pw, _ = op.best_correlation(xb[b], 4, 300, 0.01, trunc)
        assert np.array_equal(bc.periods[b], per), ("bcorr", b)
        np.testing.assert_allclose(bc.powers[b], pw, rtol=RTOL, atol=1e-15)


def test_mbest_pipelined_host_upload_matches_device_resident():
    """Large host batches are uploaded in pieces on a side stream with one launch per piece
    (pyperiod_b200/_device.py); results must equal the single-launch device-resident call."""
    import torch
    from pyperiod_b200 import Periods, _device
    B, N, hop = _device.PIPELINE_MIN_WINDOWS + 1000, 256, 64
    rng = np.random.default_rng(5)
    stream = rng.standard_normal((B - 1) * hop + N)
    P = Periods()
    dev = torch.from_numpy(stream).cuda()
    ref = P.m_best(torch.as_strided(dev, (B, N), (hop, 1)), num=3, max_length=64)
    pinned = torch.from_numpy(stream).pin_memory()
    got = P.m_best(torch.as_strided(pinned, (B, N), (hop, 1)), num=3, max_length=64)
    strided = np.lib.stride_tricks.as_strided(stream, shape=(B, N), strides=(hop * 8, 8), writeable=False)
    got_np = P.m_best(strided, num=3, max_length=64)
    dense = P.m_best(np.ascontiguousarray(strided), num=3, max_length=64)
    rp, rw, rs = ref.periods.cpu().numpy().view(np.uint32), ref.powers.cpu().numpy(), ref.status.cpu().numpy()
    for name, g in (("pinned", got), ("strided numpy", got_np), ("dense numpy", dense)):
        assert isinstance(g.periods, np.ndarray)
        bad = np.nonzero((g.periods != rp).any(axis=1) | (g.powers != rw).any(axis=1) | (g.status != rs))[0]
        assert bad.size == 0, (name, bad.size, bad[:8].tolist(), g.periods[bad[:2]].tolist(), rp[bad[:2]].tolist(),
                               g.powers[bad[:2]].tolist(), rw[bad[:2]].tolist(), g.status[bad[:4]].tolist())


def test_best_correlation_hierarchical_nomination_is_exact(P):
    """best_correlation ranks hierarchically and re-evaluates the near-maximal candidates with sequential folds
    (pp_sweep.cuh); the result must be bit-identical to the all-sequential sweep, also on inputs full of exact ties."""
    from pyperiod_b200 import _lib
    n = 2048
    rows = [synth.synth(n, 70_000 + i) for i in range(40)]
    t = np.arange(n)
    rows.append(np.sin(2 * np.pi * t / 10.0))                       # clean period 10: ties at 10, 20, 30, ...
    rows.append((t % 64 == 0).astype(np.float64))                   # impulse train: |S| equal for many periods
    rows.append(np.ones(n))                                          # constant
    rows.append(np.where(t < 300, synth.synth(n, 5)[:n], 0.0))       # mostly zero padding
    rows.append(synth.synth(n, 6) * 1e-200)                          # tiny scale
    rows.append(np.sign(np.sin(2 * np.pi * t / 37.0)))               # square wave, integer-valued sums
    xb = np.stack(rows)
    out = {}
    try:
        for mode in (_lib.FOLD_DIRECT, _lib.FOLD_HIERARCHICAL):
            _lib.set_fold_mode(mode)
            for trunc, orth in ((False, False), (True, True)):
                with warnings.catch_warnings():
                    warnings.simplefilter("ignore")
                    out[mode, trunc] = P(trunc, orth).best_correlation(xb, num=6, return_bases=True)
    finally:
        _lib.set_fold_mode(_lib.FOLD_HIERARCHICAL)
    for trunc in (False, True):
        a, b = out[_lib.FOLD_DIRECT, trunc], out[_lib.FOLD_HIERARCHICAL, trunc]
        assert np.array_equal(a.periods, b.periods)
        assert np.array_equal(a.powers, b.powers)
        assert np.array_equal(a.status, b.status)
        assert np.array_equal(a.bases, b.bases, equal_nan=True)


def test_mbest_f32_nomination_mode_is_exact(P):
    """PP_FOLD_NOMINATE_F32: the ranking sweep runs in float and only nominates; every candidate whose error bound
    reaches the best one is folded sequentially in fp64.  Period lists and sweep counts must equal the all-fp64
    direct sweep, powers to rounding (the north star's "fp32 option keeps periods exact")."""
    import torch
    from pyperiod_b200 import _lib
    B = 192
    stream = synth.synth_stream(B)
    win = torch.as_strided(torch.from_numpy(stream).cuda(), (B, 4096), (512, 1))
    small = synth.synth_batch(16, 1000, 4711)
    odd = synth.synth_batch(12, 3001, 4712)
    out = {}
    try:
        for mode in (_lib.FOLD_DIRECT, _lib.FOLD_NOMINATE_F32):
            _lib.set_fold_mode(mode)
            out[mode, "m"] = P().m_best(win, num=10, max_length=1024)
            out[mode, "g"] = P().m_best_gamma(win, num=10, max_length=1024)
            out[mode, "s"] = P().m_best(small, num=5)
            out[mode, "o"] = P().m_best_gamma(odd, num=6, min_length=3, max_length=999)
    finally:
        _lib.set_fold_mode(_lib.FOLD_HIERARCHICAL)
    for k in ("m", "g", "s", "o"):
        a, b = out[_lib.FOLD_DIRECT, k], out[_lib.FOLD_NOMINATE_F32, k]
        g = (lambda t: t.cpu().numpy() if hasattr(t, "cpu") else np.asarray(t))
        assert np.array_equal(g(a.periods), g(b.periods)), k
        assert np.array_equal(g(a.sweeps), g(b.sweeps)), k
        np.testing.assert_allclose(g(b.powers), g(a.powers), rtol=1e-12)


# ------------------------------------------------------------------ near-tie audit (exact re-ranking)
def test_mbest_exact_ties_vs_reference_fixture(P):
    """Exactly periodic integer-valued inputs: the projections onto p, 2p, 3p ... are bit-identical, the reference's
    norms tie exactly and strict '>' keeps the lowest period (Periods.py:507-515).  Energies rebuilt from reciprocal
    weights differ by an ulp there, so the device re-ranks every candidate inside the rounding bound the reference's
    way.  Fixture: tests/golden/ties.npz (generated from the reference), every fold mode, m_best and m_best_gamma."""
    from conftest import tie_inputs
    from pyperiod_b200 import _lib
    g = load_golden("ties")
    inputs = tie_inputs()
    names = sorted(inputs)
    xb = np.stack([inputs[k] for k in names])
    for mode in (_lib.FOLD_HIERARCHICAL, _lib.FOLD_DIRECT, _lib.FOLD_NOMINATE_F32, _lib.FOLD_HIERARCHICAL_NO_RIDERS):
        for tag, gamma in (("norm", False), ("gamma", True)):
            algo = P(fold_mode=mode)
            res = (algo.m_best_gamma if gamma else algo.m_best)(xb, num=1, max_length=1024, return_bases=True)
            assert res.status.tolist() == [0] * len(names)
            for i, k in enumerate(names):
                assert int(res.periods[i, 0]) == int(g[f"{k}_{tag}_period"][0]), (mode, tag, k, res.periods[i])
                np.testing.assert_allclose(res.powers[i, 0], g[f"{k}_{tag}_power"][0], rtol=1e-12)
                assert sha(res.bases[i]) == str(g[f"{k}_{tag}_base_sha"]), (mode, tag, k)
            if not gamma:
                assert (np.asarray(res.near_ties) >= 1).all()   # the multiples of the period are flagged as ties
    # 1-D call (drop-in form) on one of them
    per, pw, bs = P().m_best(inputs["binary7"], num=1, max_length=1024)
    assert per.tolist() == [7] and abs(pw[0] - 1.0) < 1e-12


def test_sweep_exact_ties_pick_the_lowest_period(P):
    from conftest import tie_inputs
    inputs = tie_inputs(2000)
    for k in ("int7", "binary10", "square8", "impulse9"):
        x = inputs[k]
        _, bp, bv = P().sweep(x, metric="norm", max_length=600)
        want = max(range(2, 601), key=lambda p: (op.periodic_norm(op.project(x, p)), -p))
        assert int(bp[0]) == want == int("".join(c for c in k if c.isdigit())), (k, int(bp[0]), want)
        np.testing.assert_allclose(bv[0], op.periodic_norm(op.project(x, want)), rtol=1e-13)


def test_near_ties_are_zero_on_noisy_windows_and_flagged_on_clean_sines(P):
    """Inputs with a noise floor never have two candidates inside the rounding bound (near_ties == 0: the exact-list
    claim covers them).  A clean sine is the excluded class: p, 2p, 3p ... agree to rounding, the reference's own pick
    among them depends on its BLAS; the device reports the tie and returns a multiple of the true period whose power
    is 1 to rounding.  With the gamma norm (energy / p) there is no tie and the fundamental wins as in the reference."""
    xb = synth.synth_batch(32, 4096, 30_000)
    res = P().m_best(xb, num=10, max_length=1024)
    assert np.asarray(res.near_ties).tolist() == [0] * 32
    n = np.arange(4096)
    for p0 in (10, 12, 25):
        x = np.sin(2 * np.pi * n / p0)
        r = P().m_best(x[None, :], num=1, max_length=1024)
        assert int(r.periods[0, 0]) % p0 == 0 and int(r.near_ties[0]) == 1
        assert abs(float(r.powers[0, 0]) - 1.0) < 1e-12
        rg = P().m_best_gamma(x[None, :], num=1, max_length=1024)
        per0, pw0, _ = op.m_best_gamma(x, 1, 1024)
        assert int(rg.periods[0, 0]) == int(per0[0]) == p0
        np.testing.assert_allclose(rg.powers[0, 0], pw0[0], rtol=1e-12)


def test_best_frequency_fused_round_cases(P):
    """pp_best_frequency_round: every window of a batch gets its own period in one launch; win_size != N; a window whose
    DC bin is the spectral peak is flagged (the reference raises OverflowError on int(round(inf)), Periods.py:386) while
    the rest of the batch is unaffected."""
    xb = synth.synth_batch(6, 1500, 4200)
    xb[4] += 3.0                                       # DC-dominated window
    res = P().best_frequency(xb, num=4)
    assert res.status.tolist() == [0, 0, 0, 0, 1, 0]
    assert len({tuple(res.periods[b].tolist()) for b in (0, 1, 2, 3, 5)}) > 1     # windows differ in their periods
    for b in (0, 1, 2, 3, 5):
        per0, pw0, bs0 = op.best_frequency(xb[b], None, 4)
        assert np.array_equal(res.periods[b], per0)
        np.testing.assert_allclose(res.powers[b], pw0, rtol=RTOL)
        np.testing.assert_allclose(res.bases[b], bs0, rtol=RTOL, atol=1e-14)
    with pytest.raises(OverflowError):
        P().best_frequency(xb[4], num=2)
    x = synth.synth(1000, 4300)
    for win in (4096, 1000, 999):
        per, pw, bs = P(True, False).best_frequency(x, win_size=win, num=3)
        per0, pw0, bs0 = op.best_frequency(x, win, 3, True, False)
        assert np.array_equal(per, per0)
        np.testing.assert_allclose(pw, pw0, rtol=RTOL)
        np.testing.assert_allclose(bs, bs0, rtol=RTOL, atol=1e-14)


def test_single_process_multi_device_sharding(P):
    """Periods(devices=[...]): one process, contiguous row blocks per device, one worker thread each.  A one-GPU box
    runs it with the same device listed twice (two threads, two workspaces, the shard / merge logic); with more GPUs
    the same call spreads over them."""
    import torch
    xb = synth.synth_batch(37, 1024, 5100)
    ndev = torch.cuda.device_count()
    devs = [0, 1 % ndev, 0][: 3]
    one = P().m_best(xb, num=5, max_length=300)
    many = P(devices=devs).m_best(xb, num=5, max_length=300)
    assert np.array_equal(one.periods, many.periods) and np.array_equal(one.powers, many.powers)
    assert np.array_equal(one.sweeps, many.sweeps) and many.status.shape == (37,)
    s1 = P().small_to_large(xb, thresh=0.08)
    s2 = P(devices=devs).small_to_large(xb, thresh=0.08)
    assert np.array_equal(s1.count, s2.count)
    for b in range(37):
        assert s1.window(b)[0] == s2.window(b)[0]
    c1 = P(True, True).best_correlation(xb, num=3, max_length=200)
    c2 = P(True, True, devices=devs).best_correlation(xb, num=3, max_length=200)
    assert np.array_equal(c1.periods, c2.periods) and np.array_equal(c1.powers, c2.powers)
    # hop-framed stream: every shard uploads its own range of the stream (with the halo)
    stream = synth.synth_stream(64, 1024, 128, 777)
    win = synth.windows_from_stream(stream, 1024, 128)
    a = P().m_best_gamma(win, num=4, max_length=256)
    b2 = P(devices=devs).m_best_gamma(win, num=4, max_length=256)
    assert np.array_equal(a.periods, b2.periods) and np.array_equal(a.powers, b2.powers)
