"""One TF32 (tcgen05) Ramanujan periodogram call for ncu: python tools/prof_ram_tf32.py [B]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pyperiod_b200 import RamanujanPeriods, synth
B = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
x = torch.from_numpy(synth.synth_batch(B, 4096, 50_000)).cuda()
r = RamanujanPeriods(precision="tf32")
for _ in range(2):
    n = r.find_periods(x)
torch.cuda.synchronize()
print("ok", float(n.sum()))
