import sys, os
sys.path.insert(0, "/root/repo")
import numpy as np, torch
from pyperiod_b200 import Periods, _lib, synth
B = 4096
stream = synth.synth_stream(B)
dev = torch.from_numpy(stream).cuda()
win = torch.as_strided(dev, (B, 4096), (512, 1))
P = Periods()
res = {}
for mode, name in ((_lib.FOLD_HIERARCHICAL, "hier"), (_lib.FOLD_NOMINATE_F32, "f32")):
    _lib.set_fold_mode(mode)
    for fn in ("m_best", "m_best_gamma"):
        r = getattr(P, fn)(win, num=10, max_length=1024)
        res[name, fn] = (r.periods.cpu().numpy(), r.powers.cpu().numpy(), r.sweeps.cpu().numpy(), r.status.cpu().numpy())
for fn in ("m_best", "m_best_gamma"):
    a, b = res["hier", fn], res["f32", fn]
    same = (a[0] == b[0]).all(axis=1)
    print(fn, "windows with identical period lists:", int(same.sum()), "of", B, "| sweeps equal", int((a[2] == b[2]).sum()),
          "| max rel power diff on identical:", float(np.max(np.abs(a[1][same] - b[1][same]) / np.abs(a[1][same]))), "| status", int((b[3] != 0).sum()))
    if not same.all():
        i = int(np.nonzero(~same)[0][0]); print(" first mismatch", i, a[0][i], b[0][i])
