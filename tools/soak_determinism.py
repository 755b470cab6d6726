"""Volume determinism probe: every batched algorithm on many short windows, twice, bit for bit (dynamic window
hand-out makes a cross-window state leak show up as a run-to-run difference).  python tools/soak_determinism.py"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from pyperiod_b200 import Periods, QOPeriods, RamanujanPeriods

def eq(a, b):
    if a is None or b is None: return a is b
    if isinstance(a, torch.Tensor): return bool(torch.equal(a, b)) if a.dtype != torch.float64 else bool(((a == b) | (a.isnan() & b.isnan())).all())
    return np.array_equal(np.asarray(a), np.asarray(b), equal_nan=True)

def check(name, fn, fields):
    r1, r2 = fn(), fn()
    bad = [f for f in fields if not eq(getattr(r1, f), getattr(r2, f))]
    print(f"{name:60s} {'ok' if not bad else 'DIFFERENT: ' + ','.join(bad)}", flush=True)
    return not bad

rng = np.random.default_rng(2024)
ok = True
for N, hop, B in ((256, 64, 20000), (600, 150, 12000), (1000, 1000, 6000)):
    stream = torch.from_numpy(rng.standard_normal((B - 1) * hop + N)).cuda()
    win = torch.as_strided(stream, (B, N), (hop, 1))
    for pmax in (17, 40, 64, 126, N // 3):
        for trunc, orth in ((False, False), (True, False), (True, True)):
            P = Periods(trunc, orth)
            tag = f"N={N} pmax={pmax} trunc={int(trunc)} orth={int(orth)}"
            ok &= check("m_best " + tag, lambda: P.m_best(win, num=4, max_length=pmax), ("periods", "powers", "status", "sweeps"))
            ok &= check("m_best_gamma " + tag, lambda: P.m_best_gamma(win, num=4, max_length=pmax), ("periods", "powers", "status"))
            ok &= check("best_correlation " + tag, lambda: P.best_correlation(win, num=4, max_length=pmax), ("periods", "powers", "status"))
        ok &= check(f"small_to_large N={N} n_periods={pmax}", lambda: Periods().small_to_large(win, thresh=0.05, n_periods=pmax), ("periods", "powers", "count", "status"))
    sub = win[:4000]
    for pmax in (40, 100, N // 3):
        ok &= check(f"QO find N={N} pmax={pmax}", lambda: QOPeriods().find_periods(sub, num=3, thresh=0.05, max_length=pmax), ("periods", "norms", "n_weights", "weights", "res", "status"))
        ok &= check(f"QO find trunc N={N} pmax={pmax}", lambda: QOPeriods(trunc_to_integer_multiple=True).find_periods(sub, num=3, thresh=0.05, max_length=pmax), ("periods", "norms", "n_weights", "weights", "res", "status"))
    ok &= check(f"Ramanujan with weights N={N}", lambda: RamanujanPeriods().find_periods_with_weights(sub[:1500], thresh=0.3, max_length=min(200, N // 3)), ("periods", "n_weights", "weights", "res", "status"))
print("ALL OK" if ok else "DIFFERENCES FOUND")
