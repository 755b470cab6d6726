"""QOPeriods.find_periods throughput + phase breakdown: python tools/perf_qo.py [B]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from pyperiod_b200 import QOPeriods, _lib, synth
B = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
base = synth.synth_batch(256, 4096, 50_000)
x = torch.from_numpy(np.concatenate([base * (1 - 0.001 * r) for r in range(-(-B // 256))])[:B].copy()).cuda()
prof = torch.zeros(8, dtype=torch.int64, device="cuda")
_lib.set_profile_buffer(prof)
q = QOPeriods()
q.find_periods(x, num=4, thresh=0.05, return_res=False); torch.cuda.synchronize(); prof.zero_()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); r = q.find_periods(x, num=4, thresh=0.05, return_res=False); e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1)
p = prof.cpu().numpy().astype(float)
print(f"B={B} {ms:.1f} ms {B/ms*1e3:.0f} win/s rows={float(r.n_weights.float().mean()):.0f} status!=0: {int((r.status!=0).sum())}")
print("cycles/window: sweep %.0f layout %.0f rhs+tables %.0f factor %.0f (gram entries %.0f, diagonal blocks %.0f) solves+recon+refine %.0f" % (p[0]/p[4], p[1]/p[4], p[2]/p[4], p[3]/p[4], p[6]/p[4], p[7]/p[4], p[5]/p[4]))
rows = r.n_weights.double()
flop = float((rows ** 3 / 3).sum())
print("factor flop/window (R^3/3 of the final dictionary) %.1f M; %.2f TFLOP/s over the whole call" % (flop / B / 1e6, flop / ms / 1e9))
