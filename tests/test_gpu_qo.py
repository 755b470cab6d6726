"""GPU parity tests for QOPeriods.find_periods / get_periods (default branch) against the golden
fixtures generated from the reference and against the oracle."""
import numpy as np
import pytest

from conftest import load_golden
from oracle import qo as oq
from pyperiod_b200 import synth

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def QO():
    from pyperiod_b200 import QOPeriods
    return QOPeriods


def _check(d, res, g, pre):
    assert np.array_equal(np.asarray(d["periods"]), g[pre + "periods"])
    np.testing.assert_allclose(d["norms"], g[pre + "norms"], rtol=1e-10)
    assert [int(k) for k in d["basis_dictionary"]] == g[pre + "dict_keys"].tolist()
    assert list(d["basis_dictionary"].values()) == g[pre + "dict_vals"].tolist()
    np.testing.assert_allclose(d["weights"], g[pre + "weights"], rtol=1e-8, atol=1e-11)
    np.testing.assert_allclose(res, g[pre + "res"], rtol=0, atol=1e-11)
    assert d["subspaces"].shape == (len(d["weights"]), len(res))


def test_qo_readme_vs_golden(QO):
    g = load_golden("readme_qo_ram")
    c = synth.readme_signal(0)
    d, res = QO().find_periods(c, num=2, thresh=0.05)
    _check(d, res, g, "qo_")
    assert d["periods"].tolist() == [59, 100] and d["basis_dictionary"] == {"59": 59, "100": 99}
    d3, res3 = QO().find_periods(c, num=3, thresh=0.05)
    _check(d3, res3, g, "qo3_")
    # reconstruction from the returned dictionary reproduces data - res
    np.testing.assert_allclose(d3["subspaces"].T @ d3["weights"], c - res3, rtol=0, atol=1e-12)


def test_qo_synth_4096_vs_golden_and_batch(QO):
    g = load_golden("qo_ram_synth")
    xb = synth.synth_batch(2, 4096, 50_000)
    out = QO().find_periods(xb, num=4, thresh=0.05)
    assert out.status.tolist() == [0, 0]
    for b in range(2):
        d, res = out.window(b)
        _check(d, res, g, f"qo_{b}_")
        d1, res1 = QO().find_periods(xb[b], num=4, thresh=0.05)      # 1-D call == batch row
        assert np.array_equal(d1["weights"], d["weights"]) and np.array_equal(res1, res)


def test_qo_vs_oracle_modes(QO):
    xb = synth.synth_batch(3, 1500, 8800)
    for trunc in (False, True):
        for num, thresh in ((1, 0.5), (3, 0.05), (4, 0.6)):
            out = QO(trunc_to_integer_multiple=trunc).find_periods(xb, num=num, thresh=thresh, max_length=300)
            for b in range(3):
                d, res = out.window(b)
                d0, res0 = oq.find_periods(xb[b], num=num, thresh=thresh, max_length=300, trunc=trunc)
                assert np.array_equal(np.asarray(d["periods"]), np.asarray(d0["periods"])), (trunc, num, b)
                assert d["basis_dictionary"] == d0["basis_dictionary"]
                np.testing.assert_allclose(d["norms"], d0["norms"], rtol=1e-10)
                np.testing.assert_allclose(d["weights"], d0["weights"], rtol=1e-8, atol=1e-11)
                np.testing.assert_allclose(res, res0, rtol=0, atol=1e-11)


def test_qo_zero_input(QO):
    d, res = QO().find_periods(np.zeros(600), num=2, thresh=0.05)
    assert d["periods"].tolist() == [1] and d["basis_dictionary"] == {"1": 600} and not res.any()


def test_get_periods_vs_golden(QO):
    g = load_golden("readme_qo_ram")
    c = synth.readme_signal(0)
    q = QO()
    d, _ = q.find_periods(c, num=2, thresh=0.05)
    gp = q.get_periods(d["weights"], d["basis_dictionary"], "lstsq")
    for i, v in enumerate(gp):
        np.testing.assert_allclose(v, g[f"qo_getp_lstsq_{i}"], rtol=0, atol=1e-9)
    d3, _ = q.find_periods(c, num=3, thresh=0.05)
    for kind, key in (("row reduction", "rowreduction"), ("lstsq", "lstsq")):
        if f"qo3_getp_{key}_linalgerror" in g.files:
            continue
        gp = q.get_periods(d3["weights"], d3["basis_dictionary"], kind)
        for i, v in enumerate(gp):
            np.testing.assert_allclose(v, g[f"qo3_getp_{key}_{i}"], rtol=0, atol=1e-8)


def test_unsupported_branches_raise(QO):
    x = synth.synth(512, 1)
    with pytest.raises(NotImplementedError):
        QO(orthogonalize=True).find_periods(x, num=1, thresh=0.1)
    with pytest.raises(NotImplementedError):
        QO().find_periods(x, num=1, thresh=0.1, update_weights=False)
    with pytest.raises(ValueError):
        QO(basis_type="fourier").find_periods(x, num=1, thresh=0.1)
    with pytest.raises(TypeError):
        QO().find_periods(x, num=2)


def test_qo_gcds_extracted_vs_golden():
    """QOPeriodsWithGCDsExtracted: device loop + host layout + device re-solve against the
    reference's own outputs (tests/golden/qo_gcd.npz), 1-D and batched."""
    from pyperiod_b200 import QOPeriodsWithGCDsExtracted as QG
    g = load_golden("qo_gcd")
    for i in range(4):
        seed, n, num, max_length = (int(v) for v in g[f"c{i}_args"])
        x = synth.synth(n, seed)
        d, res = QG().find_periods(x, num=num, thresh=float(g[f"c{i}_thresh"]), max_length=max_length or None)
        assert np.array_equal(np.asarray(d["periods"]), g[f"c{i}_periods"])
        assert [int(k) for k in d["basis_dictionary"]] == g[f"c{i}_dict_keys"].tolist()
        assert [int(v) for v in d["basis_dictionary"].values()] == g[f"c{i}_dict_vals"].tolist()
        np.testing.assert_allclose(d["norms"], g[f"c{i}_norms"], rtol=1e-10)
        np.testing.assert_allclose(d["weights"], g[f"c{i}_weights"], rtol=1e-8, atol=1e-11)
        np.testing.assert_allclose(res, g[f"c{i}_res"], rtol=0, atol=1e-11)
        assert d["subspaces"].shape == (len(d["weights"]), n)
    # batch of two equal-length windows == the 1-D calls
    xb = np.stack([synth.synth(1500, 8800), synth.synth(1500, 8801)])
    out = QG().find_periods(xb, num=3, thresh=0.05, max_length=300)
    for b in range(2):
        d1, r1 = QG().find_periods(xb[b], num=3, thresh=0.05, max_length=300)
        db, rb = out.window(b)
        assert db["basis_dictionary"] == d1["basis_dictionary"]
        assert np.array_equal(db["weights"], d1["weights"]) and np.array_equal(rb, r1)


# ------------------------------------------------------------------ dictionaries of any size (no row cap)
def _weights_close(got, want, gram_cond, what):
    """Normwise agreement of two fp64 solutions of the same normal equations: both carry a forward error of
    about cond(G) * eps, so that is the floor; 1e-10 relative is asked wherever the conditioning allows it."""
    err = np.max(np.abs(got - want)) / np.max(np.abs(want))
    tol = max(1e-10, 50 * gram_cond * 2.2e-16)
    assert err <= tol, (what, err, tol, gram_cond)


def test_qo_weights_normwise_1e10(QO):
    """North-star tolerance: weights within 1e-10 (relative to the largest weight) of the reference's LU solution on
    the golden config-5 QO windows (cond(G) of a few thousand), residuals to 1e-12."""
    g = load_golden("qo_ram_synth")
    xb = synth.synth_batch(2, 4096, 50_000)
    out = QO().find_periods(xb, num=4, thresh=0.05)
    for b in range(2):
        d, res = out.window(b)
        w0 = g[f"qo_{b}_weights"]
        assert np.max(np.abs(d["weights"] - w0)) / np.max(np.abs(w0)) <= 1e-10
        np.testing.assert_allclose(res, g[f"qo_{b}_res"], rtol=0, atol=1e-12)


@pytest.mark.parametrize("how", ["pool", "rerun", "pool_exhausted"])
def test_qo_second_launch_for_large_dictionaries(QO, monkeypatch, how):
    """Windows whose dictionary outgrows the dense weights array keep their weights in the overflow pool of the first
    launch ("pool"); those that outgrow the factor storage, or find the pool exhausted, are re-run with room for the
    rows they need ("rerun", "pool_exhausted"): same results as a single launch that had the room from the start, and
    as the oracle."""
    from pyperiod_b200 import qoperiods
    xb = synth.synth_batch(6, 1500, 8800)
    ref = QO().find_periods(xb, num=4, thresh=0.05, max_length=400)
    monkeypatch.setattr(qoperiods, "RMAX_FIRST", 96)
    if how == "rerun":
        monkeypatch.setattr(qoperiods, "RMAX_FACTOR", 128)
    if how == "pool_exhausted":
        monkeypatch.setattr(qoperiods, "POOL_MIN_SLOTS", 1)
        monkeypatch.setattr(qoperiods, "POOL_FRACTION", 1000)
    out = QO().find_periods(xb, num=4, thresh=0.05, max_length=400)
    assert out.big is not None and len(out.big) >= 1
    assert out.status.tolist() == ref.status.tolist() == [0] * 6
    for b in range(6):
        d, res = out.window(b)
        d1, res1 = ref.window(b)
        assert np.array_equal(np.asarray(d["periods"]), np.asarray(d1["periods"]))
        assert d["basis_dictionary"] == d1["basis_dictionary"]
        np.testing.assert_allclose(d["weights"], d1["weights"], rtol=0, atol=1e-12 * np.max(np.abs(d1["weights"])))
        np.testing.assert_allclose(res, res1, rtol=0, atol=1e-13)
        d0, res0 = oq.find_periods(xb[b], num=4, thresh=0.05, max_length=400)
        assert np.array_equal(np.asarray(d["periods"]), np.asarray(d0["periods"]))
        np.testing.assert_allclose(res, res0, rtol=0, atol=1e-11)


def test_qo_pipelined_download_matches_device_resident(QO, monkeypatch):
    """A large host batch is uploaded in pieces and its residuals / weights come back piece by piece through the
    two staging buffers of RowDownloader (here shrunk to 1 MB so that every piece takes several chunks, with the
    threaded host copy forced on): same results as the device-resident call."""
    import torch
    from pyperiod_b200 import _device
    monkeypatch.setattr(_device, "PIPELINE_MIN_WINDOWS", 512)
    monkeypatch.setattr(_device, "STAGE_BYTES", 1 << 20)
    monkeypatch.setattr(_device, "COPY_THREADS_MIN_BYTES", 1 << 16)
    _device._stage_bufs.clear()                      # staging buffers are sized when first used
    rng = np.random.default_rng(11)
    base = synth.synth_batch(64, 512, 9100)
    xb = np.concatenate([base * (1.0 + 0.01 * k) + 1e-3 * rng.standard_normal(base.shape) for k in range(10)])[:600]
    ref = QO().find_periods(torch.from_numpy(xb).cuda(), num=3, thresh=0.05, max_length=150)
    got = QO().find_periods(torch.from_numpy(xb).pin_memory(), num=3, thresh=0.05, max_length=150)
    assert isinstance(got.res, np.ndarray) and isinstance(got.weights, np.ndarray)
    assert np.array_equal(got.res, ref.res.cpu().numpy())
    assert np.array_equal(got.weights, ref.weights.cpu().numpy())
    assert np.array_equal(np.asarray(got.periods), ref.periods.cpu().numpy().view(np.asarray(got.periods).dtype))
    assert np.array_equal(np.asarray(got.status), ref.status.cpu().numpy())
    _device._stage_bufs.clear()


def test_to_host_staged_path(monkeypatch):
    """to_host above STAGED_D2H_MIN_BYTES: chunks through the page-locked staging buffers into an ordinary array
    (odd sizes, several chunks, int and float payloads)."""
    import torch
    from pyperiod_b200 import _device
    monkeypatch.setattr(_device, "STAGED_D2H_MIN_BYTES", 1 << 20)
    monkeypatch.setattr(_device, "PINNED_D2H_MIN_BYTES", 1 << 10)
    monkeypatch.setattr(_device, "STAGE_BYTES", 1 << 20)
    monkeypatch.setattr(_device, "COPY_THREADS_MIN_BYTES", 1 << 16)
    _device._stage_bufs.clear()
    g = torch.Generator(device="cuda").manual_seed(3)
    for shape, dt in (((1237, 1031), torch.float64), ((5_000_011,), torch.int32), ((3, 700, 1001), torch.float64)):
        t = (torch.rand(shape, device="cuda", generator=g) * 1000).to(dt)
        h = _device.to_host(t)
        assert isinstance(h, np.ndarray) and h.shape == tuple(shape)
        assert np.array_equal(h, t.cpu().numpy())
    tt = torch.arange(4_000_000, device="cuda", dtype=torch.float64).reshape(2000, 2000).t()   # non-contiguous
    assert np.array_equal(_device.to_host(tt), tt.cpu().numpy())
    _device._stage_bufs.clear()


def test_qo_rows_beyond_samples_are_singular(QO):
    """More dictionary rows than samples: A A^T is singular by rank.  The reference's LU either raises LinAlgError or
    returns rounding noise; the device reports SINGULAR and keeps the previous round's outputs (QOPeriods.py:552-559)."""
    x = synth.synth(240, 77)
    out = QO().find_periods(x[None, :], num=6, thresh=0.0, max_length=119)
    st = int(out.status[0])
    assert st in (0, 3)
    if st == 3:
        assert int(out.n_weights[0]) <= 240


# ------------------------------------------------------------------ basis_type="ramanujan", get_periods on the device
RAMBASIS_CASES = [(600, 5, 2, None), (600, 5, 3, None), (1024, 50_001, 3, None), (2000, 7, 4, 300)]
GETP_CASES = [(2000, 7, 4, 300), (1024, 50_001, 3, 200), (4096, 50_003, 4, None)]


def test_qo_ramanujan_basis_vs_golden(QO):
    """QOPeriods(basis_type="ramanujan").find_periods (QOPeriods.py:970-971, 1005-1052) against the reference's own
    outputs: periods, dictionary, norms and residual.  The reference's Gram matrix is singular by construction there
    (q shifted rows span phi(q) dimensions), its weights are one arbitrary solution; the device returns the
    minimum-norm one, which must reconstruct the same signal."""
    g = load_golden("qo_rambasis")
    for i, (n, seed, num, ml) in enumerate(RAMBASIS_CASES):
        x = synth.synth(n, seed)
        d, res = QO(basis_type="ramanujan").find_periods(x, num=num, thresh=0.05, max_length=ml)
        assert np.asarray(d["periods"]).tolist() == g[f"c{i}_periods"].tolist()
        assert [int(k) for k in d["basis_dictionary"]] == g[f"c{i}_dict_keys"].tolist()
        assert [int(v) for v in d["basis_dictionary"].values()] == g[f"c{i}_dict_vals"].tolist()
        np.testing.assert_allclose(d["norms"], g[f"c{i}_norms"], rtol=1e-10)
        np.testing.assert_allclose(res, g[f"c{i}_res"], rtol=0, atol=1e-10)
        a = d["subspaces"]
        assert a.shape == (len(d["weights"]), n)
        np.testing.assert_allclose(x - a.T @ d["weights"], res, rtol=0, atol=1e-10)
        # minimum-norm weights: no larger than the reference's
        assert np.linalg.norm(d["weights"]) <= np.linalg.norm(g[f"c{i}_weights"]) * (1 + 1e-9)
    # batch form, two windows of one shape
    xb = np.stack([synth.synth(600, 5), synth.synth(600, 6)])
    out = QO(basis_type="ramanujan").find_periods(xb, num=2, thresh=0.05)
    d0, res0 = out.window(0)
    assert np.asarray(d0["periods"]).tolist() == g["c0_periods"].tolist()
    np.testing.assert_allclose(res0, g["c0_res"], rtol=0, atol=1e-10)
    d1, res1 = out.window(1)
    import contextlib, io
    with contextlib.redirect_stdout(io.StringIO()):
        e1, r1 = oq.find_periods(xb[1], num=2, thresh=0.05, basis="ramanujan")
    assert np.asarray(d1["periods"]).tolist() == np.asarray(e1["periods"]).tolist()
    np.testing.assert_allclose(res1, r1, rtol=0, atol=1e-10)


def test_get_periods_device_vs_golden(QO):
    """QOPeriods.get_periods on the device (pp_qo_get_periods) against the reference's "row reduction" and lstsq
    branches on 3-4 period dictionaries, reference form and batch form."""
    g = load_golden("qo_rambasis")
    for i in range(len(GETP_CASES)):
        layout = {str(int(k)): int(v) for k, v in zip(g[f"g{i}_dict_keys"], g[f"g{i}_dict_vals"])}
        for kind, key in (("row reduction", "rowreduction"), ("lstsq", "lstsq"), ("qr", "lstsq")):
            gp = QO().get_periods(g[f"g{i}_weights"], layout, kind)
            assert [len(v) for v in gp] == [int(k) for k in layout]
            np.testing.assert_allclose(np.concatenate(gp), g[f"g{i}_{key}"], rtol=0, atol=1e-10)
    # batch form: the result of a batched find_periods goes straight back in
    xb = np.stack([synth.synth(1024, 50_001 + j) for j in range(5)])
    r = QO().find_periods(xb, num=3, thresh=0.01, max_length=200)
    pb = QO().get_periods(r)
    assert np.asarray(pb.status).tolist() == [0] * 5
    for b in range(5):
        d, _ = r.window(b)
        want = oq.get_periods(d["weights"], d["basis_dictionary"], "lstsq")
        got = pb.window(b)
        assert len(got) == len(want)
        for u, v in zip(got, want):
            np.testing.assert_allclose(u, v, rtol=0, atol=1e-10)


def test_get_periods_rank_one_raises_like_the_reference(QO):
    """reduce_rows returns a 1-D array when the pairwise-GCD matrix has rank one and np.linalg.solve rejects it
    (QOPeriods.py:86-94, 794); lstsq goes through."""
    w = np.arange(1.0, 13.0)
    with pytest.raises(np.linalg.LinAlgError):
        QO().get_periods(w[:5], {"5": 5})
    with pytest.raises(np.linalg.LinAlgError):
        QO().get_periods(w, {"5": 5, "7": 7})
    gp = QO().get_periods(w, {"5": 5, "7": 7}, "lstsq")
    want = oq.get_periods(w, {"5": 5, "7": 7}, "lstsq")
    for u, v in zip(gp, want):
        np.testing.assert_allclose(u, v, rtol=0, atol=1e-12)
    gp = QO().get_periods(w[:5], {"5": 5}, "lstsq")
    np.testing.assert_allclose(gp[0], w[:5] - w[:5].mean(), rtol=0, atol=1e-13)


def test_custom_test_function(QO):
    """test_function(self, data, reconstruction) (QOPeriods.py:388-391, 418): the default written out by the caller
    gives the default path's result; another rule is checked against the oracle running the same callable."""
    import contextlib, io
    x = synth.synth(1024, 50_002)
    rms = lambda v: np.sqrt(np.sum(np.power(v, 2)) / len(v))
    d0, res0 = QO().find_periods(x, num=4, thresh=0.05)
    d1, res1 = QO().find_periods(x, num=4, thresh=0.05, test_function=lambda s, a, y: rms(y) > rms(a) * 0.05)
    assert np.asarray(d0["periods"]).tolist() == np.asarray(d1["periods"]).tolist()
    assert d0["basis_dictionary"] == d1["basis_dictionary"]
    np.testing.assert_allclose(d1["weights"], d0["weights"], rtol=0, atol=1e-10 * np.max(np.abs(d0["weights"])))
    np.testing.assert_allclose(res1, res0, rtol=0, atol=1e-12)
    calls = []
    stop_after_two = lambda s, a, y: (calls.append(1) or len(calls) < 2)
    d2, res2 = QO().find_periods(x, num=6, thresh=None, test_function=stop_after_two)
    calls2 = []
    with contextlib.redirect_stdout(io.StringIO()):
        e2, r2 = oq.find_periods(x, num=6, test_function=lambda s, a, y: (calls2.append(1) or len(calls2) < 2))
    assert np.asarray(d2["periods"]).tolist() == np.asarray(e2["periods"]).tolist()
    assert d2["basis_dictionary"] == e2["basis_dictionary"]
    np.testing.assert_allclose(res2, r2, rtol=0, atol=1e-11)


def test_muresan_eq3_finder_vs_golden(QO):
    """get_best_period_orthogonal / eq_3 (QOPeriods.py:1122-1232) on the device against the
    reference's own outputs: equation-3 powers with the divisors' powers removed, plain and normalised, the selected
    period, and eq_3 itself; 1-D and batched."""
    g = load_golden("muresan")
    cases = [(600, 5, None), (1024, 50_001, 300), (2000, 7, None)]
    for i, (n, seed, max_p) in enumerate(cases):
        x = synth.synth(n, seed)
        q = QO()
        for norm in (False, True):
            want = g[f"c{i}_pows_{int(norm)}"]
            got = q.get_best_period_orthogonal(x, max_p, norm, True)
            assert got.shape == want.shape
            np.testing.assert_allclose(got, want, rtol=0, atol=1e-10 * np.max(want))
            assert q.get_best_period_orthogonal(x, max_p, norm) == int(g[f"c{i}_best_{int(norm)}"])
        got = [q.eq_3(x, p) for p in (1, 2, 7, 30, 97, n // 4)]
        np.testing.assert_allclose(got, g[f"c{i}_eq3"], rtol=1e-11)
    xb = np.stack([synth.synth(600, 5), synth.synth(600, 6), synth.synth(600, 7)])
    best = QO().get_best_period_orthogonal(xb, None, True)
    pw = QO().get_best_period_orthogonal(xb, None, True, True)
    assert pw.shape == (3, 300) and int(best[0]) == int(g["c0_best_1"])
    for b in range(3):
        np.testing.assert_allclose(pw[b], oq.get_best_period_orthogonal(xb[b], None, True, True), rtol=0,
                                   atol=1e-10 * np.max(pw[b]))
    np.testing.assert_allclose(QO().auto_corr(xb[0], 17), oq.auto_corr(xb[0], 17), rtol=1e-12)
