"""best_correlation: hierarchical nomination + exact verification against the all-sequential sweep (config-4 shape)."""
import sys, os, warnings
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from pyperiod_b200 import Periods, _lib, synth
B = int(sys.argv[1]) if len(sys.argv) > 1 else 384
x = torch.from_numpy(synth.synth_batch(B, 8192, 40_000)).cuda()
warnings.simplefilter("ignore")
for trunc, orth in ((True, True), (False, False)):
    res = {}
    for mode in (_lib.FOLD_DIRECT, _lib.FOLD_HIERARCHICAL):
        _lib.set_fold_mode(mode)
        r = Periods(trunc, orth).best_correlation(x, num=10)
        res[mode] = (r.periods.cpu().numpy(), r.powers.cpu().numpy(), r.status.cpu().numpy())
    a, b = res[_lib.FOLD_DIRECT], res[_lib.FOLD_HIERARCHICAL]
    print("trunc/orth", trunc, orth, "| windows with identical periods:", int((a[0] == b[0]).all(axis=1).sum()), "of", B,
          "| powers bit-identical:", bool(np.array_equal(a[1], b[1])), "| status equal:", bool(np.array_equal(a[2], b[2])))
_lib.set_fold_mode(_lib.FOLD_HIERARCHICAL)
