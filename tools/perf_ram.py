"""Ramanujan periodogram at config 5's shape (N = 4096, q = 2..1365): throughput, DMMA fraction and agreement of the
precision modes.  python tools/perf_ram.py [B] [modes: fp64,tf32,f32_compat]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from pyperiod_b200 import RamanujanPeriods, synth
from pyperiod_b200._device import stage_windows
B = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
modes = sys.argv[2].split(",") if len(sys.argv) > 2 else ["fp64"]
base = synth.synth_batch(min(B, 256), 4096, 50_000)
x = torch.from_numpy(np.concatenate([base * (1 - 0.001 * r) for r in range(-(-B // base.shape[0]))])[:B].copy()).cuda()
flop = 2.0 * sum(q * q for q in range(2, 1366))          # per window: one q x q circulant product per period
ref = None
for mode in modes:
    r = RamanujanPeriods(precision=mode)
    ms = []
    for rep in range(4):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        win = stage_windows(x, None)
        e0.record(); nr = r._norms_device(win, 2, 1365); e1.record(); torch.cuda.synchronize()
        ms.append(e0.elapsed_time(e1))
    best = min(ms[1:])
    line = f"{mode}: B={B} ms {['%.1f' % m for m in ms]}  {B / best * 1e3:.0f} windows/s  {flop * B / best / 1e9:.2f} TFLOP/s"
    if ref is None:
        ref = nr
    else:
        line += f"  max |d| / max norm vs {modes[0]}: {float(((nr - ref).abs().amax(1) / ref.amax(1)).max()):.2e}"
    print(line, flush=True)
