"""Where does a full-size QO call spend its time: first launch (rmax 1024) vs the re-run of oversized windows."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from pyperiod_b200 import QOPeriods, synth, _lib
B = int(sys.argv[1]) if len(sys.argv) > 1 else 32768
x = synth.synth_batch_device(B, 4096, 50_000, torch.device("cuda"))
q = QOPeriods()
def timed(fn, reps=3):
    fn(); torch.cuda.synchronize()
    out = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); r = fn(); e1.record(); torch.cuda.synchronize(); out.append(e0.elapsed_time(e1))
    return out, r
t, r = timed(lambda: q.find_periods(x, num=4, thresh=0.05, return_res=False, rmax=1024))
big = (r.status == _lib.STATUS_TOO_LARGE)
print(f"B={B} first launch only (rmax=1024): {t} ms; too_large {int(big.sum())}; rows of too_large: {sorted(r.n_weights[big].tolist())[-10:]}")
t2, r2 = timed(lambda: q.find_periods(x, num=4, thresh=0.05, return_res=False))
print(f"full call (with re-run): {t2} ms; rows max {int(r2.n_weights.max())}")
idx = torch.nonzero(big).flatten()
xb = x[idx].contiguous()
for rmax in (2048, 4096):
    t3, r3 = timed(lambda: q.find_periods(xb, num=4, thresh=0.05, return_res=False, rmax=rmax))
    print(f"re-run alone, {xb.shape[0]} windows, rmax={rmax}: {t3} ms")
t4, _ = timed(lambda: q.find_periods(x[:B // 2], num=4, thresh=0.05, return_res=False, rmax=1024))
print(f"half batch first launch only: {t4} ms")
