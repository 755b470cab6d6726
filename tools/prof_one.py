"""One M-best launch for ncu: python tools/prof_one.py [B] [gamma]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pyperiod_b200 import Periods, synth
B = int(sys.argv[1]) if len(sys.argv) > 1 else 2072
stream = synth.synth_stream(B)
dev = torch.from_numpy(stream).cuda()
win = torch.as_strided(dev, (B, 4096), (512, 1))
P = Periods()
fn = P.m_best_gamma if len(sys.argv) > 2 else P.m_best
for _ in range(2):
    r = fn(win, num=10, max_length=1024)
torch.cuda.synchronize()
print("ok", int(r.sweeps.sum()))
