"""`RamanujanPeriods` -- drop-in for pyPeriod.RamanujanPeriods (pyPeriod/RamanujanPeriods.py:61-169).

`find_periods` is the Ramanujan periodogram: a fold of every period followed by a dense contraction
with the circulant of Ramanujan sums on the FP64 tensor cores (csrc/pp_ramanujan.cu);
`find_periods_with_weights` thresholds it on the device and hands the selected periods to the
QOPeriods solve stage (csrc/pp_qo.cu).  Values are fp64; the reference stores its projection in
float32 (RamanujanPeriods.py:127), so agreement with it is bounded by ~1e-7 relative on the norms.
"""
from __future__ import annotations

import numpy as np
import torch

from . import _lib
from ._device import Workspace, ptr, stage_windows, stream_ptr
from .periods import _export
from .qoperiods import QOBatchResult, QOPeriods
from .tables import get_tables

TILE_WINDOWS = 2048  # windows whose folds are held at once (7.5 MB per window at qmax = 1365)


class RamanujanPeriods(QOPeriods):
    def __init__(self, basis_type="natural", device=None, precision="fp64"):
        """precision: "fp64" (FP64 tensor cores, default) or "tf32" (split TF32 tensor-core contraction, fp32
        accumulation: norms to ~1e-5 relative; not part of the reference API)."""
        super().__init__(basis_type, False, False, device=device)
        if precision not in ("fp64", "tf32"):
            raise ValueError("precision must be 'fp64' or 'tf32'")
        self._precision = precision
        self._verbose = None
        self._k = 0

    # ------------------------------------------------------------------ periodogram
    def _norms_device(self, w, min_length, max_length):
        lib = _lib.load()
        tb = get_tables(max_length)
        mu, phi = tb.mu_device(w.device), tb.phi_device(w.device)
        tile = min(TILE_WINDOWS, max(4, w.b))
        ws = Workspace.get(w.device, lib.pp_ramanujan_workspace_bytes(w.n, min_length, max_length, tile))
        norms = torch.zeros((w.b, max_length + 1), dtype=torch.float64, device=w.device)
        fn = lib.pp_ramanujan_norms_tf32 if self._precision == "tf32" else lib.pp_ramanujan_norms
        _lib.check(fn(ptr(w.tensor), w.ldx, w.b, w.n, int(min_length), int(max_length), ptr(mu),
                      ptr(phi), tb.pmax, tile, ptr(norms), max_length + 1, ptr(ws), ws.numel(),
                      stream_ptr(w.device)), "pp_ramanujan_norms")
        return norms

    def find_periods(self, x, min_length=2, max_length=None, select_periods=None):
        """Periodogram norms, length max_length+1, zero below min_length (RamanujanPeriods.py:67-86)."""
        w = stage_windows(x, self._device)
        if not max_length:
            max_length = w.n // 3
        norms = _export(w, self._norms_device(w, min_length, max_length))
        norms = norms[0] if w.was_1d else norms
        if select_periods:
            if hasattr(select_periods, "__call__"):
                return select_periods(norms)
            return None  # the reference falls through without a return value here
        return norms

    # ------------------------------------------------------------------ periodogram + quadratic program
    def find_periods_with_weights(self, x, min_length=2, max_length=None, thresh=0.2, kmax=32, rmax=None,
                                  return_res=True, **kwargs):
        """RamanujanPeriods.py:88-122: periods over thresh * max norm -> dictionary -> normal equations."""
        if "test_function" in kwargs:
            raise NotImplementedError("custom test_function is not evaluated on the device")
        if kwargs:
            raise TypeError(f"unexpected arguments {sorted(kwargs)}")
        lib = _lib.load()
        w = stage_windows(x, self._device)
        n, dev = w.n, w.device
        if not max_length:
            max_length = n // 3
        norms = self._norms_device(w, min_length, max_length)
        i32 = dict(dtype=torch.int32, device=dev)
        while True:
            periods = torch.zeros((w.b, kmax), **i32)
            nper = torch.zeros((w.b,), **i32)
            _lib.check(lib.pp_ramanujan_select(ptr(norms), w.b, max_length + 1, max_length + 1, float(thresh), kmax,
                                               ptr(periods), ptr(nper), stream_ptr(dev)), "pp_ramanujan_select")
            need = int(nper.max()) if w.b else 0
            if need <= kmax:
                break
            kmax = need
        if rmax is None:
            rmax = min(n, 1024)
        tb = get_tables(max_length)
        phi = tb.phi_device(dev)
        ws = Workspace.get(dev, lib.pp_qo_workspace_bytes(n, max_length, kmax, rmax))
        dict_q, dict_keep = torch.zeros((w.b, kmax), **i32), torch.zeros((w.b, kmax), **i32)
        n_dict, n_weights, status = (torch.zeros((w.b,), **i32) for _ in range(3))
        weights = torch.zeros((w.b, (rmax + 1) & ~1), dtype=torch.float64, device=dev)
        res = torch.empty((w.b, n), dtype=torch.float64, device=dev) if return_res else None
        _lib.check(lib.pp_qo_solve(ptr(w.tensor), w.ldx, w.b, n, kmax, ptr(periods), ptr(nper), int(max_length),
                                   ptr(phi), tb.pmax, int(rmax), ptr(dict_q), ptr(dict_keep), ptr(n_dict),
                                   ptr(n_weights), ptr(weights), ptr(res), ptr(status), ptr(ws), ws.numel(),
                                   stream_ptr(dev)), "pp_qo_solve")
        sel_norms = torch.gather(norms, 1, periods.long())  # norms[periods] (:116), padded entries unused
        out = QOBatchResult(_export(w, periods), _export(w, sel_norms), _export(w, nper), _export(w, dict_q),
                            _export(w, dict_keep), _export(w, n_dict), _export(w, weights), _export(w, n_weights),
                            _export(w, res), _export(w, status), n=n)
        if w.was_1d:
            st = int(out.status[0])
            if st == _lib.STATUS_SINGULAR:
                raise np.linalg.LinAlgError("Singular matrix")  # the reference's np.linalg.solve raises here
            if st == _lib.STATUS_TOO_LARGE:
                raise ValueError("dictionary has more rows than rmax; pass a larger rmax")
            d, r = out.window(0)
            d["periods"] = np.asarray(d["periods"]).astype(np.int64)  # np.argwhere indices (:97-101)
            self._output = d
            return d, r
        return out
