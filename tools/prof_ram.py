"""One Ramanujan periodogram call for ncu: python tools/prof_ram.py [B] [qmax]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pyperiod_b200 import RamanujanPeriods, synth
B = int(sys.argv[1]) if len(sys.argv) > 1 else 512
qmax = int(sys.argv[2]) if len(sys.argv) > 2 else 1365
x = torch.from_numpy(synth.synth_batch(B, 4096, 50_000)).cuda()
r = RamanujanPeriods()
for _ in range(2):
    n = r.find_periods(x, 2, qmax)
torch.cuda.synchronize()
print("ok", float(n.sum()))
