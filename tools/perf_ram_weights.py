"""RamanujanPeriods.find_periods_with_weights at config 5's shape: python tools/perf_ram_weights.py [B]"""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from pyperiod_b200 import RamanujanPeriods, synth
B = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
base = synth.synth_batch(min(B, 256), 4096, 50_000)
x = torch.from_numpy(np.concatenate([base * (1 - 0.001 * r) for r in range(-(-B // base.shape[0]))])[:B].copy()).cuda()
r = RamanujanPeriods()
for rep in range(2):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    nr = r.find_periods(x); torch.cuda.synchronize(); t1 = time.perf_counter()
    out = r.find_periods_with_weights(x, thresh=0.2, return_res=False); torch.cuda.synchronize(); t2 = time.perf_counter()
    rows = out.n_weights.double()
    print(f"rep {rep}: B={B} norms {t1-t0:.3f}s ({B/(t1-t0):.0f} win/s)  norms+select+solve {t2-t1:.3f}s ({B/(t2-t1):.0f} win/s)"
          f"  status!=0: {int((out.status != 0).sum())}  rows mean {float(rows.mean()):.0f} max {int(rows.max())}"
          f"  >1024: {int((rows > 1024).sum())}  factor GFLOP total {float((rows**3/3).sum())/1e9:.1f}")
    st = out.status.cpu().numpy(); rw = out.n_weights.cpu().numpy()
    print("   status histogram", {int(k): int((st == k).sum()) for k in np.unique(st)}, " rows of non-ok windows", sorted(rw[st != 0].tolist())[:40])
