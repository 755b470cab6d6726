"""One best-correlation launch for ncu: python tools/prof_bcorr.py [B]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from pyperiod_b200 import Periods, synth
B = int(sys.argv[1]) if len(sys.argv) > 1 else 592
base = synth.synth_batch(min(B, 64), 8192, 40_000)
x = torch.from_numpy(np.concatenate([base * (1 - 0.001 * r) for r in range(-(-B // base.shape[0]))])[:B].copy()).cuda()
P = Periods(True, True)
for _ in range(2):
    r = P.best_correlation(x, num=10)
torch.cuda.synchronize()
print("ok", int((r.status == 0).sum()))
