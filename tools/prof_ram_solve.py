"""One find_periods_with_weights call at config 5's shape for ncu (the solve kernels): python tools/prof_ram_solve.py [B]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pyperiod_b200 import RamanujanPeriods, synth
B = int(sys.argv[1]) if len(sys.argv) > 1 else 512
x = torch.from_numpy(synth.synth_batch(B, 4096, 50_000)).cuda()
r = RamanujanPeriods()
out = r.find_periods_with_weights(x, thresh=0.2, return_res=False)
torch.cuda.synchronize()
print("ok rows mean", float(out.n_weights.double().mean()), "max", int(out.n_weights.max()))
