"""Load woolgathering/pyPeriod from /root/reference with the SURVEY.md §8c patch set.

BUILD-CONTAINER ONLY.  /root/reference does not exist on the GPU box, so nothing
under tests/ imports this module at test time; it is used by make_golden.py (and
by oracle validation runs done by hand) to generate the committed fixtures.

The reference at HEAD does not import (QOPeriods.py:86 uses an un-imported `Any`)
and six further one-line defects break orthogonalize / m_best / QOPeriods /
RamanujanPeriods.  Each substitution below restores evident intent and changes no
arithmetic; each asserts its exact match count so drift in the reference is loud.
Sources are copied to a temp dir, never into this repository.
"""
from __future__ import annotations

import importlib
import os
import shutil
import sys
import tempfile

REFERENCE_ROOT = "/root/reference"

# (file, old, new, expected_count)
PATCHES = [
    # P1  Periods.py:209  non-existent kwarg in orthogonalize branch
    ("Periods.py", "get_factors(p, remove_1_and_n=True)",
     "get_factors(p, remove_1=True, remove_n=True)", 1),
    # P2  Periods.py:548  same, in M-best step 2
    ("Periods.py", "get_factors(periods[i], remove_1_and_n=True)",
     "get_factors(periods[i], remove_1=True, remove_n=True)", 1),
    # P3  QOPeriods.py:86  `Any` never imported
    ("QOPeriods.py", "import warnings\n", "import warnings\nfrom typing import Any\n", 1),
    # P4  QOPeriods.py:148  class uses Periods' methods but has no base
    ("QOPeriods.py", "class QOPeriods:\n", "class QOPeriods(Periods):\n", 1),
    # P5  QOPeriods.py:190  wrong class in super()
    ("QOPeriods.py", "super(Periods, self).__init__(trunc_to_integer_multiple, orthogonalize)",
     "super(QOPeriods, self).__init__(trunc_to_integer_multiple, orthogonalize)", 1),
    # P6  QOPeriods.py:726-734  `_k` passed into positional `type`
    ("QOPeriods.py", "self._k, type=", "k=self._k, type=", 4),
    # P7a RamanujanPeriods.py:62-65  __init__ never sets _k
    ("RamanujanPeriods.py", "        self._verbose = None\n",
     "        self._verbose = None\n        self._k = 0\n", 1),
    # P7b RamanujanPeriods.py:109  unpack order opposite to solve_quadratic's return
    ("RamanujanPeriods.py", "resconst, output_weights = self.solve_quadratic(",
     "output_weights, resconst = self.solve_quadratic(", 1),
]


def load(reference_root: str = REFERENCE_ROOT):
    """Return the patched `pyPeriod` package module."""
    src = os.path.join(reference_root, "pyPeriod")
    if not os.path.isdir(src):
        raise RuntimeError(f"{src} not found: the reference only exists in the build container")
    tmp = tempfile.mkdtemp(prefix="pyperiod_ref_")
    dst = os.path.join(tmp, "pyPeriod")
    shutil.copytree(src, dst)
    os.chmod(dst, 0o755)
    for f in os.listdir(dst):
        os.chmod(os.path.join(dst, f), 0o644)
    dup = os.path.join(dst, "periods.py")  # byte-identical duplicate of Periods.py
    if os.path.exists(dup):
        os.remove(dup)
    for fname, old, new, count in PATCHES:
        path = os.path.join(dst, fname)
        with open(path) as fh:
            text = fh.read()
        got = text.count(old)
        assert got == count, f"{fname}: expected {count}x {old!r}, found {got}"
        with open(path, "w") as fh:
            fh.write(text.replace(old, new))
    sys.path.insert(0, tmp)
    for name in [m for m in sys.modules if m == "pyPeriod" or m.startswith("pyPeriod.")]:
        del sys.modules[name]
    return importlib.import_module("pyPeriod")
