"""Debug aid: rerun tests/test_gpu_qo.py::test_qo_vs_oracle_modes and print every mismatching weight."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from pyperiod_b200 import QOPeriods as QO, synth
from oracle import qo as oq

xb = synth.synth_batch(3, 1500, 8800)
for rep in range(3):
    for trunc in (False, True):
        for num, thresh in ((1, 0.5), (3, 0.05), (4, 0.6)):
            out = QO(trunc_to_integer_multiple=trunc).find_periods(xb, num=num, thresh=thresh, max_length=300)
            for b in range(3):
                d, res = out.window(b)
                d0, res0 = oq.find_periods(xb[b], num=num, thresh=thresh, max_length=300, trunc=trunc)
                w, w0 = np.asarray(d["weights"]), np.asarray(d0["weights"])
                if w.shape != w0.shape:
                    print(rep, trunc, num, b, "shape", w.shape, w0.shape, d["periods"], d0["periods"]); continue
                bad = np.nonzero(~np.isclose(w, w0, rtol=1e-8, atol=1e-11))[0]
                if bad.size:
                    print(rep, trunc, num, thresh, b, "periods", d["periods"], d["basis_dictionary"], "bad idx", bad[:10],
                          "got", w[bad[:5]], "want", w0[bad[:5]], "ratio", w[bad[:5]] / w0[bad[:5]], flush=True)
print("done")
