"""Where the end-to-end time of QOPeriods.find_periods on 65,536 host windows goes: python tools/probe_qo_e2e.py [B]"""
import sys, os, time, cProfile, pstats
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pyperiod_b200 import QOPeriods, synth
B = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
x = synth.synth_batch_device(B, 4096, 50_000, torch.device("cuda:0"))
xh = torch.empty((B, 4096), dtype=torch.float64, pin_memory=True); xh.copy_(x); torch.cuda.synchronize()
q = QOPeriods()
for rep in range(4):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    out = q.find_periods(xh, num=4, thresh=0.05)
    torch.cuda.synchronize(); print(f"rep {rep}: e2e {1e3*(time.perf_counter()-t0):.0f} ms", flush=True)
pr = cProfile.Profile(); pr.enable()
out = q.find_periods(xh, num=4, thresh=0.05); torch.cuda.synchronize()
pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(18)
