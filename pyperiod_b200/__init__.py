"""pyperiod_b200 -- B200-native drop-in for the projection hot path of woolgathering/pyPeriod.

    from pyperiod_b200 import Periods            # same surface as pyPeriod.Periods (+ (B, N) batches)

All arithmetic runs in hand-written sm_100a CUDA kernels (csrc/) behind the C ABI of
include/pyperiod_b200.h; there is no CPU path.  `synth` holds the deterministic input
generators used by the tests and the benchmark.
"""
from . import synth  # noqa: F401  (numpy only)

__all__ = ["Periods", "BatchResult", "QOPeriods", "QOBatchResult", "QOPeriodsWithGCDsExtracted", "RamanujanPeriods",
           "synth"]


def __getattr__(name):
    # product classes import torch and load the CUDA library; keep `import pyperiod_b200.synth` light
    if name in ("Periods", "BatchResult"):
        from . import periods as _p
        return getattr(_p, name)
    if name in ("QOPeriods", "QOBatchResult", "QOPeriodsWithGCDsExtracted", "QOGcdBatchResult"):
        from . import qoperiods as _q
        return getattr(_q, name)
    if name == "RamanujanPeriods":
        from . import ramanujan as _r
        return _r.RamanujanPeriods
    raise AttributeError(name)
