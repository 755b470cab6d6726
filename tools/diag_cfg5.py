"""Accuracy of the Ramanujan + QP solve on ill-conditioned config-5 dictionaries: GPU (Cholesky on the tensor cores +
refinement) and numpy's LU (the reference's np.linalg.solve), both against a solution refined in extended precision."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, scipy.linalg as sl
from oracle import qo as oq
from pyperiod_b200 import RamanujanPeriods, synth

wins = (6, 14, 18, 4, 16)
xb = np.stack([synth.synth(4096, 50_000 + b) for b in wins])
for refine in (0, 1, 3):
    out = RamanujanPeriods().find_periods_with_weights(xb, thresh=0.2, refine=refine)
    for i, b in enumerate(wins):
        d, res = out.window(i)
        a, _ = oq.get_subspaces(np.asarray(d["periods"]), 4096)
        G = a @ a.T
        lu = sl.lu_factor(G)
        w_lu = sl.lu_solve(lu, a @ xb[i])
        Gl, al = G.astype(np.longdouble), a.astype(np.longdouble)
        bl = al @ xb[i].astype(np.longdouble)
        w = w_lu.astype(np.longdouble)
        for _ in range(8):
            w = w + sl.lu_solve(lu, (bl - Gl @ w).astype(np.float64))
        res_star = xb[i].astype(np.longdouble) - al.T @ w
        wn = float(np.abs(w).max())
        print(f"refine={refine} window {b}: R={a.shape[0]} cond={np.linalg.cond(G):.2e} |w|max={wn:.3g}  "
              f"weights err gpu {float(np.abs(d['weights'] - w).max()) / wn:.2e} lu {float(np.abs(w_lu - w).max()) / wn:.2e}  "
              f"residual err gpu {float(np.abs(res - res_star).max()):.2e} lu {float(np.abs(xb[i] - a.T @ w_lu - res_star).max()):.2e}",
              flush=True)
