"""Per-call times of m_best with and without the bases streamed out."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from pyperiod_b200 import Periods, synth
B = 8192
dev = torch.from_numpy(synth.synth_stream(B)).cuda()
win = torch.as_strided(dev, (B, 4096), (512, 1))
P = Periods()
for bases in (False, True, False, True):
    ts = []
    for _ in range(6):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        r = P.m_best(win, num=10, max_length=1024, return_bases=bases)
        torch.cuda.synchronize(); ts.append((time.perf_counter() - t0) * 1e3)
    print("bases" if bases else "plain", " ".join(f"{t:7.1f}" for t in ts), flush=True)
