"""QOPeriods.find_periods on distinct config-5 windows (65,536 by default): wall time and the launches it takes.
python tools/probe_qo_full.py [B]"""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pyperiod_b200 import QOPeriods, synth
B = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
x = synth.synth_batch_device(B, 4096, 50_000, torch.device("cuda:0"))
q = QOPeriods()
q.find_periods(x[:4096], num=4, thresh=0.05)
for rep in range(3):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    out = q.find_periods(x, num=4, thresh=0.05)
    torch.cuda.synchronize(); t1 = time.perf_counter()
    rows = out.n_weights
    print(f"rep {rep}: B={B} {1e3*(t1-t0):.1f} ms  {B/(t1-t0):.0f} windows/s  rows mean {float(rows.float().mean()):.0f} max {int(rows.max())} "
          f">1024: {int((rows>1024).sum())}  status!=0: {int((out.status!=0).sum())}  pooled/re-run windows: {0 if out.big is None else len(out.big)}", flush=True)
