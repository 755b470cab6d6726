"""One small_to_large launch at config 2's shape for ncu: python tools/prof_s2l.py [B]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pyperiod_b200 import Periods, synth
B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
x = synth.synth_batch_device(B, 2048, 20_000, torch.device("cuda:0"))
P = Periods()
for _ in range(2):
    r = P.small_to_large(x, thresh=0.1)
torch.cuda.synchronize()
print("ok", float(r.count.float().mean()))
