"""How the 2 GB of residuals of a 65,536-window QO batch best reach the host: python tools/probe_d2h.py"""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
n = 65536 * 4096
src = torch.rand(n, dtype=torch.float64, device="cuda")
torch.cuda.synchronize()
def t(fn, label, reps=3):
    out = []
    for _ in range(reps):
        torch.cuda.synchronize(); t0 = time.perf_counter(); r = fn(); torch.cuda.synchronize(); out.append(time.perf_counter() - t0); del r
    print(f"{label:60s} " + " ".join(f"{v*1e3:7.1f}" for v in out) + " ms", flush=True)
t(lambda: torch.empty(n, dtype=torch.float64, pin_memory=True), "pinned alloc 2.1 GB (freed each time: caching host allocator)")
keep = []
def alloc_keep():
    h = torch.empty(n, dtype=torch.float64, pin_memory=True); keep.append(h); return None
t(alloc_keep, "pinned alloc 2.1 GB (kept alive: fresh cudaHostAlloc)", reps=2)
h = keep[0]
t(lambda: h.copy_(src, non_blocking=True), "D2H into an existing pinned buffer")
t(lambda: src.cpu(), "src.cpu() (pageable)")
stage = [torch.empty(8 << 20, dtype=torch.float64, pin_memory=True) for _ in range(2)]   # 2 x 64 MB
cs = torch.cuda.Stream()
def staged():
    out = np.empty(n, dtype=np.float64)
    ch = stage[0].numel()
    evs = [None, None]
    k = 0
    pending = []
    for o in range(0, n, ch):
        m = min(ch, n - o)
        b = k & 1
        if evs[b] is not None:
            evs[b].synchronize()
            po, pm = pending[b]
            np.copyto(out[po:po + pm], stage[b][:pm].numpy())
        with torch.cuda.stream(cs):
            stage[b][:m].copy_(src[o:o + m], non_blocking=True)
            ev = torch.cuda.Event(); ev.record(cs)
        evs[b] = ev
        if len(pending) < 2: pending.append((o, m))
        else: pending[b] = (o, m)
        k += 1
    for b in range(2):
        if evs[b] is not None:
            evs[b].synchronize(); po, pm = pending[b]; np.copyto(out[po:po + pm], stage[b][:pm].numpy())
    return out
t(staged, "staged through 2 x 64 MB pinned buffers into fresh pageable numpy")
