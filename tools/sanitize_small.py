"""Small invocation of every kernel for `compute-sanitizer --tool racecheck|memcheck python tools/sanitize_small.py`."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from pyperiod_b200 import Periods, QOPeriods, RamanujanPeriods, synth

which = sys.argv[1].split(",") if len(sys.argv) > 1 else ["periods", "qo", "ram"]
xb = synth.synth_batch(3, 600, 777)
if "periods" in which:
    for trunc, orth in ((False, False), (True, True)):
        P = Periods(trunc_to_integer_multiple=trunc, orthogonalize=orth)
        r = P.m_best(xb, num=4, max_length=128)
        r = P.m_best_gamma(xb, num=4, max_length=128)
        r = P.small_to_large(xb, thresh=0.1, n_periods=100)
        r = P.best_correlation(xb, num=3, max_length=100)
    print("periods ok", flush=True)
if "qo" in which:
    out = QOPeriods().find_periods(xb, num=3, thresh=0.05, max_length=100)
    out = QOPeriods(trunc_to_integer_multiple=True).find_periods(xb, num=2, thresh=0.05, max_length=100)
    print("qo ok", flush=True)
if "periods" in which:
    P = Periods(trunc_to_integer_multiple=True, orthogonalize=False)     # truncated hierarchical sweep
    r = P.m_best(xb, num=4, max_length=128)
    r = Periods().best_frequency(xb, num=2)
    print("periods (trunc hierarchy, best_frequency) ok", flush=True)
if "ram" in which:
    R = RamanujanPeriods()
    out = R.find_periods_with_weights(xb, max_length=60, thresh=0.2)
    print("ram ok", flush=True)
    for prec in ("tf32", "f32_compat"):
        n = RamanujanPeriods(precision=prec).find_periods(xb, 2, 60)
    print("ram tf32 / f32_compat ok", flush=True)
if "qo" in which:
    q = QOPeriods(basis_type="ramanujan")
    out = q.find_periods(xb, num=2, thresh=0.05, max_length=60)
    q2 = QOPeriods()
    d, res = q2.find_periods(xb[0], num=3, thresh=0.05, max_length=100)
    gp = q2.get_periods(d["weights"], d["basis_dictionary"])
    pw = q2.get_best_period_orthogonal(xb[0], None, True, True)
    print("qo ramanujan basis / get_periods / muresan ok", flush=True)
