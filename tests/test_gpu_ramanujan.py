"""GPU parity tests for RamanujanPeriods: periodogram (dense DMMA contraction on folded sums) and
find_periods_with_weights, against the golden fixtures of the reference and the fp64 closed form."""
import numpy as np
import pytest

from conftest import load_golden
from oracle import ramanujan as oram
from pyperiod_b200 import synth

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def R():
    from pyperiod_b200 import RamanujanPeriods
    return RamanujanPeriods


def test_norms_vs_fp64_closed_form(R):
    for n, seed, qmax in ((1024, 50_000, None), (2000, 7, 150), (999, 3, 333)):
        x = synth.synth(n, seed)
        got = R().find_periods(x, 2, qmax)
        want = oram.norms_closed_form_f64(x, 2, qmax)
        assert got.shape == want.shape and got[0] == 0 and got[1] == 0
        np.testing.assert_allclose(got[2:], want[2:], rtol=1e-11, atol=0)


def test_norms_vs_reference_float32_limited(R):
    g = load_golden("readme_qo_ram")
    c = synth.readme_signal(0)
    got = R().find_periods(c, 2, 120)
    np.testing.assert_allclose(got, g["ram_norms"], rtol=2e-6)   # the reference stores float32 (RamanujanPeriods.py:127)
    g2 = load_golden("qo_ram_synth")
    xb = synth.synth_batch(2, 1024, 50_000)
    nb = R().find_periods(xb)
    for b in range(2):
        np.testing.assert_allclose(nb[b], g2[f"ram_{b}_norms"], rtol=2e-6)


def test_batch_tiles_and_odd_sizes(R):
    xb = synth.synth_batch(7, 600, 123)          # 7 windows: partial 4-window fold group and partial GEMM tile
    nb = R().find_periods(xb, 3, 200)
    for b in range(7):
        want = oram.norms_closed_form_f64(xb[b], 3, 200)
        np.testing.assert_allclose(nb[b, 3:], want[3:], rtol=1e-11)
        assert not nb[b, :3].any()


def test_find_periods_with_weights_vs_golden(R):
    g = load_golden("readme_qo_ram")
    c = synth.readme_signal(0)
    d, res = R().find_periods_with_weights(c, max_length=120, thresh=0.2)
    assert d["periods"].tolist() == g["ram_periods"].tolist() == [58, 60, 100, 102]
    assert [int(k) for k in d["basis_dictionary"]] == g["ram_dict_keys"].tolist()
    assert list(d["basis_dictionary"].values()) == g["ram_dict_vals"].tolist()
    np.testing.assert_allclose(d["norms"], g["ram_sel_norms"], rtol=2e-6)
    np.testing.assert_allclose(d["weights"], g["ram_weights"], rtol=1e-8, atol=1e-11)
    np.testing.assert_allclose(res, g["ram_res"], rtol=0, atol=1e-11)
    g2 = load_golden("qo_ram_synth")
    xb = synth.synth_batch(2, 1024, 50_000)
    out = R().find_periods_with_weights(xb, thresh=0.2)
    for b in range(2):
        d, res = out.window(b)
        assert np.asarray(d["periods"]).tolist() == g2[f"ram_{b}_periods"].tolist()
        assert list(d["basis_dictionary"].values()) == g2[f"ram_{b}_dict_vals"].tolist()
        np.testing.assert_allclose(d["weights"], g2[f"ram_{b}_weights"], rtol=1e-7, atol=1e-10)
        np.testing.assert_allclose(res, g2[f"ram_{b}_res"], rtol=0, atol=1e-10)


def test_tf32_option_tracks_fp64():
    """precision="tf32": split-TF32 tensor-core contraction (fp32 accumulate); norms within 1e-5 of the fp64 path
    relative to the largest norm, and the thresholded period selection is unchanged."""
    from pyperiod_b200 import RamanujanPeriods
    xb = synth.synth_batch(70, 2048, 61_000)          # 70 windows: a full 64-window tile and a ragged one
    a = RamanujanPeriods().find_periods(xb, 2, 400)
    b = RamanujanPeriods(precision="tf32").find_periods(xb, 2, 400)
    scale = a.max(axis=1, keepdims=True)
    assert np.max(np.abs(a - b) / scale) < 1e-5
    assert np.all(b[:, :2] == 0)
    ra = RamanujanPeriods().find_periods_with_weights(xb[:8], max_length=200, thresh=0.2)
    rb = RamanujanPeriods(precision="tf32").find_periods_with_weights(xb[:8], max_length=200, thresh=0.2)
    for i in range(8):
        if int(ra.status[i]) == 0 and int(rb.status[i]) == 0:
            da, _ = ra.window(i)
            db, _ = rb.window(i)
            assert np.array_equal(np.asarray(da["periods"]), np.asarray(db["periods"]))
