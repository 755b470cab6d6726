"""Throughput of the secondary BASELINE.json configs (2, 4, 5) on one GPU, device-resident inputs.
Not the contract bench (bench.py is config 3); prints one JSON object per line.

    python tools/bench_configs.py [--scale 1.0]      scale < 1 shrinks the batch sizes
"""
import argparse, json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from pyperiod_b200 import Periods, QOPeriods, RamanujanPeriods, synth


def timed(fn, reps=2):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        out = fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e-3, out


def batch(b, n, seed0):
    # distinct windows without generating b of them on the host: 256 generated, tiled with per-copy scaling
    base = synth.synth_batch(min(b, 256), n, seed0)
    reps = -(-b // base.shape[0])
    x = np.concatenate([base * (1.0 - 0.001 * r) for r in range(reps)])[:b]
    return torch.from_numpy(np.ascontiguousarray(x)).cuda()


ap = argparse.ArgumentParser(); ap.add_argument("--scale", type=float, default=1.0); args = ap.parse_args()
S = args.scale
# config 2: small_to_large(thresh=0.1), N=2048, 16,384 windows
B = int(16384 * S); x = batch(B, 2048, 20_000)
t, r = timed(lambda: Periods().small_to_large(x, thresh=0.1))
print(json.dumps({"config": 2, "algo": "small_to_large(thresh=0.1)", "N": 2048, "windows": B, "seconds": t,
                  "windows_per_s": B / t, "mean_periods": float(r.count.float().mean())}), flush=True)
# config 4: Muresan-Parks best_correlation(num=10), N=8192 (256K windows in the config; a slice here)
B = int(8192 * S); x = batch(B, 8192, 40_000)
t, r = timed(lambda: Periods(True, True).best_correlation(x, num=10), reps=2)
adds = 10 * 2728 * 8192
print(json.dumps({"config": 4, "algo": "Periods(True,True).best_correlation(num=10)", "N": 8192, "windows": B,
                  "seconds": t, "windows_per_s": B / t, "smem_GBps_algorithmic": B * adds * 8 / t / 1e9}), flush=True)
# config 5a: QOPeriods.find_periods(num=4, thresh=0.05), N=4096 (65,536 windows in the config; a slice here)
B = int(8192 * S); x = batch(B, 4096, 50_000)
t, r = timed(lambda: QOPeriods().find_periods(x, num=4, thresh=0.05, return_res=False), reps=2)
print(json.dumps({"config": "5-QO", "algo": "QOPeriods.find_periods(num=4, thresh=0.05)", "N": 4096, "windows": B,
                  "seconds": t, "windows_per_s": B / t, "mean_rows": float(r.n_weights.float().mean()),
                  "status_nonzero": int((r.status != 0).sum())}), flush=True)
# config 5b: Ramanujan periodogram q = 2..1365, N=4096
B = int(2048 * S); x = batch(B, 4096, 50_000)
t, r = timed(lambda: RamanujanPeriods().find_periods(x), reps=2)
flops = 2.0 * sum(q * q for q in range(2, 1366)) * B
print(json.dumps({"config": "5-Ramanujan", "algo": "RamanujanPeriods.find_periods (dense DMMA contraction)", "N": 4096,
                  "qmax": 1365, "windows": B, "seconds": t, "windows_per_s": B / t, "dmma_TFLOPs": flops / t / 1e12,
                  "flops_per_window": flops / B}), flush=True)
t, r = timed(lambda: RamanujanPeriods().find_periods_with_weights(x, thresh=0.2, return_res=False), reps=2)
print(json.dumps({"config": "5-Ramanujan+QP", "algo": "RamanujanPeriods.find_periods_with_weights(thresh=0.2)",
                  "N": 4096, "windows": B, "seconds": t, "windows_per_s": B / t,
                  "status_nonzero": int((r.status != 0).sum()), "mean_rows": float(r.n_weights.float().mean())}), flush=True)
