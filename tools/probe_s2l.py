import sys, time
sys.path.insert(0, "/root/repo")
import numpy as np, torch
from pyperiod_b200 import Periods, synth
base = synth.synth_batch(256, 2048, 20_000)
x = torch.from_numpy(np.concatenate([base * (1.0 - 0.001 * r) for r in range(64)])[:16384].copy()).cuda()
P = Periods()
for rep in range(3):
    ts = []
    for _ in range(8):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        r = P.small_to_large(x, thresh=0.1)
        torch.cuda.synchronize(); ts.append((time.perf_counter() - t0) * 1e3)
    print("s2l ms per call:", " ".join(f"{t:6.1f}" for t in ts), " mean periods", float(r.count.float().mean()))
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10): r = P.small_to_large(x, thresh=0.1)
e1.record(); torch.cuda.synchronize()
print("event ms per call", e0.elapsed_time(e1) / 10)
