"""e2e timing probe: per-step times of Periods().m_best on a pinned host stream view (the bench's e2e leg)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from pyperiod_b200 import Periods, synth

B = int(sys.argv[1]) if len(sys.argv) > 1 else 131072
stream = synth.synth_stream(B)
host = torch.from_numpy(stream).pin_memory()
hw = torch.as_strided(host, (B, 4096), (512, 1))
dev = host.cuda()
dw = torch.as_strided(dev, (B, 4096), (512, 1))
P = Periods()
for name, w in (("device", dw), ("host-pinned", hw), ("device", dw), ("host-pinned", hw)):
    P.m_best(w, num=10, max_length=1024)
    torch.cuda.synchronize()
    ts = []
    for _ in range(4):
        t0 = time.perf_counter()
        r = P.m_best(w, num=10, max_length=1024)
        torch.cuda.synchronize()
        ts.append((time.perf_counter() - t0) * 1e3)
    print(f"{name:12s} ms/step " + " ".join(f"{t:8.1f}" for t in ts) + f"   best {B / min(ts) * 1e3:9.0f} win/s", flush=True)
