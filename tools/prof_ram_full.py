"""One full-size config-5 call (65,536 windows) for a launch list: python tools/prof_ram_full.py [B]"""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pyperiod_b200 import RamanujanPeriods, synth
B = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
x = synth.synth_batch_device(B, 4096, 50_000, torch.device("cuda:0"))
r = RamanujanPeriods()
out = r.find_periods_with_weights(x[:2048], thresh=0.2)     # warm-up on a slice
torch.cuda.synchronize(); t0 = time.perf_counter()
out = r.find_periods_with_weights(x, thresh=0.2)
torch.cuda.synchronize(); t1 = time.perf_counter()
rows = out.n_weights.double()
print(f"B={B} wall {t1-t0:.3f}s  {B/(t1-t0):.0f} windows/s  rows mean {float(rows.mean()):.0f} max {int(rows.max())}  "
      f">1024: {int((rows>1024).sum())}  >3328: {int((rows>3328).sum())}  status!=0: {int((out.status!=0).sum())}")
h = torch.histc(rows.float(), bins=16, min=0, max=4096)
print("rows histogram (256-wide bins):", [int(v) for v in h])
fl = (rows ** 3 / 3)
print("factor GFLOP: total %.0f, in windows >1024 rows %.0f, >2048 %.0f, >3328 %.0f" % (float(fl.sum())/1e9, float(fl[rows>1024].sum())/1e9, float(fl[rows>2048].sum())/1e9, float(fl[rows>3328].sum())/1e9))
