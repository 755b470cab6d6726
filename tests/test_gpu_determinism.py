"""Run-to-run determinism of every kernel (GPU).

The persistent kernels hand out work dynamically (tops to warps, windows to CTAs) and share on-chip
scratch between phases; a missing barrier shows up as a result that changes between identical calls
(compute-sanitizer is not available on the GPU pool).  Each algorithm is run several times on the
same batch and every output must be bit-identical."""
import numpy as np
import pytest

from pyperiod_b200 import synth

pytestmark = pytest.mark.gpu

REPS = 4


def _same(a, b):
    if a is None and b is None:
        return True
    a, b = np.asarray(a), np.asarray(b)
    return a.shape == b.shape and np.array_equal(a, b, equal_nan=True)


@pytest.mark.parametrize("trunc,orth", [(False, False), (True, True)])
def test_periods_algorithms_repeatable(trunc, orth):
    from pyperiod_b200 import Periods
    xb = synth.synth_batch(700, 1024, 91_000)
    P = Periods(trunc_to_integer_multiple=trunc, orthogonalize=orth)
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        calls = {
            "m_best": lambda: P.m_best(xb, num=6, max_length=300),
            "m_best_gamma": lambda: P.m_best_gamma(xb, num=6, max_length=300),
            "small_to_large": lambda: P.small_to_large(xb, thresh=0.05, n_periods=256),
            "best_correlation": lambda: P.best_correlation(xb, num=4, max_length=256),
        }
        for name, fn in calls.items():
            first = fn()
            for _ in range(REPS - 1):
                again = fn()
                assert _same(first.periods, again.periods), name
                assert _same(first.powers, again.powers), name
                assert _same(first.status, again.status), name


def test_qo_repeatable():
    from pyperiod_b200 import QOPeriods
    xb = synth.synth_batch(600, 1500, 92_000)
    q = QOPeriods()
    first = q.find_periods(xb, num=3, thresh=0.05, max_length=300)
    for _ in range(REPS - 1):
        again = q.find_periods(xb, num=3, thresh=0.05, max_length=300)
        for b in range(0, 600, 7):
            d0, r0 = first.window(b)
            d1, r1 = again.window(b)
            assert _same(d0["periods"], d1["periods"]) and _same(d0["weights"], d1["weights"]), b
            assert _same(d0["norms"], d1["norms"]) and _same(r0, r1), b


def test_ramanujan_repeatable():
    from pyperiod_b200 import RamanujanPeriods
    xb = synth.synth_batch(300, 1024, 93_000)
    r = RamanujanPeriods()
    first = r.find_periods(xb, max_length=128)
    for _ in range(REPS - 1):
        assert _same(first, r.find_periods(xb, max_length=128))


@pytest.mark.parametrize("pmax", [24, 64, 100, 126])
def test_mbest_small_max_length_repeatable_and_exact(pmax):
    """max_length < 127: the per-warp scratch of M-best step 2 (8 x 32 doubles) is larger than the two single-period
    vectors it shares shared memory with; the plan has to size that region for it.  (It did not: about one window in
    two million came out with garbage powers at max_length = 64, depending on which window a CTA had processed
    before.)  Many short windows, twice, bit for bit -- and a sample against the oracle."""
    import torch
    from oracle import periods as op
    from pyperiod_b200 import Periods
    B, N, hop = 20_000, 256, 64
    rng = np.random.default_rng(100 + pmax)
    stream = torch.from_numpy(rng.standard_normal((B - 1) * hop + N)).cuda()
    win = torch.as_strided(stream, (B, N), (hop, 1))
    a = Periods().m_best(win, num=3, max_length=pmax)
    b = Periods().m_best(win, num=3, max_length=pmax)
    assert torch.equal(a.periods.view(torch.int32), b.periods.view(torch.int32)) and torch.equal(a.powers, b.powers)
    assert float(a.powers.min()) >= 0.0 and bool(torch.isfinite(a.powers).all())
    host = stream.cpu().numpy()
    for i in (0, 1, 777, 12_345, B - 1):
        per0, pw0, _ = op.m_best(host[i * hop: i * hop + N], 3, pmax)
        assert a.periods[i].cpu().numpy().view(np.uint32).tolist() == [int(v) for v in per0]
        np.testing.assert_allclose(a.powers[i].cpu().numpy(), pw0, rtol=1e-10)
