"""Solve-stage time of ONE large config-5 dictionary replicated over the SMs: python tools/probe_solve_big.py [seed_offset] [copies]"""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from pyperiod_b200 import RamanujanPeriods, synth
off = int(sys.argv[1]) if len(sys.argv) > 1 else 14
copies = int(sys.argv[2]) if len(sys.argv) > 2 else 148
x1 = synth.synth(4096, 50_000 + off)
r = RamanujanPeriods()
norms = r.find_periods(x1)
per = np.argwhere(norms / np.abs(np.max(norms)) > 0.2).flatten()
x = torch.from_numpy(np.tile(x1, (copies, 1))).cuda()
pick = lambda n: per          # same periods for every copy: only the solve stage differs from nothing
for rep in range(3):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    nr = r.find_periods(x); torch.cuda.synchronize(); t1 = time.perf_counter()
    out = r.find_periods_with_weights(x, thresh=0.2, return_res=False); torch.cuda.synchronize(); t2 = time.perf_counter()
    R = int(out.n_weights[0])
    solve = (t2 - t1) - (t1 - t0)
    print(f"rep {rep}: window {off} x{copies}: R={R} periods={len(per)} solve {solve*1e3:.1f} ms  "
          f"-> {R**3/3*copies/solve/1e12:.2f} TFLOP/s, {solve*1.965e9/ -(-copies//148)/1e6:.1f} Mcycles per window per SM", flush=True)
