import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from pyperiod_b200 import QOPeriods, synth
B = int(sys.argv[1]) if len(sys.argv) > 1 else 592
base = synth.synth_batch(min(B, 256), 4096, 50_000)
x = torch.from_numpy(np.concatenate([base * (1 - 0.001 * r) for r in range(-(-B // base.shape[0]))])[:B].copy()).cuda()
q = QOPeriods()
for _ in range(2):
    r = q.find_periods(x, num=4, thresh=0.05, return_res=False)
torch.cuda.synchronize(); print("ok", int(r.n_weights.sum()))
