"""Why do two timing loops of the same QO call disagree?  python tools/probe_qo_timing.py [B]"""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from pyperiod_b200 import QOPeriods, synth
B = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
base = synth.synth_batch(256, 4096, 50_000)
x = torch.from_numpy(np.concatenate([base * (1 - 0.001 * r) for r in range(-(-B // 256))])[:B].copy()).cuda()
q = QOPeriods()
for rep in range(6):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); r = q.find_periods(x, num=4, thresh=0.05, return_res=False); e1.record(); torch.cuda.synchronize()
    print(f"rep {rep}: events {e0.elapsed_time(e1):.1f} ms  wall {(time.perf_counter()-t0)*1e3:.1f} ms  too_large-first-pass? rows max {int(r.n_weights.max())}")
# distinct windows instead of 256 tiled
xd = torch.from_numpy(synth.synth_batch(B, 4096, 50_000)).cuda()
for rep in range(4):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); r = q.find_periods(xd, num=4, thresh=0.05, return_res=False); e1.record(); torch.cuda.synchronize()
    print(f"distinct rep {rep}: events {e0.elapsed_time(e1):.1f} ms rows mean {float(r.n_weights.float().mean()):.0f} max {int(r.n_weights.max())} status!=0 {int((r.status!=0).sum())}")
