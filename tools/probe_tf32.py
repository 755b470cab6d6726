"""tcgen05 TF32 periodogram against the fp64 DMMA path: accuracy and time.  python tools/probe_tf32.py [B] [N] [qmax]"""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from pyperiod_b200 import RamanujanPeriods, synth
B = int(sys.argv[1]) if len(sys.argv) > 1 else 300
N = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
qmax = int(sys.argv[3]) if len(sys.argv) > 3 else N // 3
x = synth.synth_batch_device(B, N, 50_000, torch.device("cuda"))
def run(prec):
    r = RamanujanPeriods(precision=prec)
    out = r.find_periods(x, 2, qmax); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); out = r.find_periods(x, 2, qmax); e1.record(); torch.cuda.synchronize()
    return out, e0.elapsed_time(e1)
a, ta = run("fp64")
b, tb = run("tf32")
scale = a.amax(dim=1, keepdim=True)
err = ((a - b).abs() / scale).amax().item()
rel = ((a - b).abs() / a.abs().clamp_min(1e-300))[:, 2:].amax().item()
print(f"B={B} N={N} qmax={qmax}: fp64 {ta:.1f} ms ({B/ta*1e3:.0f} win/s)  tf32 {tb:.1f} ms ({B/tb*1e3:.0f} win/s)  "
      f"max |diff| / max norm = {err:.3e}  max rel = {rel:.3e}  nan: {int(torch.isnan(b).sum())}")
sel = lambda t: [(t[i] / t[i].max() > 0.2).nonzero().flatten().tolist() for i in range(min(B, 64))]
print("thresholded period lists equal on", sum(int(u == v) for u, v in zip(sel(a), sel(b))), "of", min(B, 64))
