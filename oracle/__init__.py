"""CPU oracle for the pyPeriod projection hot path.  TEST INFRASTRUCTURE, NOT PRODUCT.

This package is a numpy restatement of the algorithms of woolgathering/pyPeriod
on the path BASELINE.json's north star names.  Only
`tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline / `--impl
reference` legs may import it, and only as the checker or the timed CPU
baseline.  Nothing under `pyperiod_b200/` imports it; the product fails loudly
when its CUDA library is missing.

Parity pin: the reference ships no tests, golden vectors or fixtures
(tests/context.py:1-7 is a sys.path shim), so the pin is the reference itself,
run in the build container with the 7-substitution patch set of SURVEY.md §8c
(tests/golden/patched_reference.py) by tests/golden/make_golden.py; its outputs
are committed as tests/golden/*.npz and tests/test_oracle_golden.py checks this
oracle against them (bit-exact for projections, exact period lists, <=1e-12
relative for norms that go through BLAS).

Modules
-------
numtheory  : factor sets in CPython set order, primes, Euler phi
periods    : project / periodic_norm / small_to_large / best_correlation / M-best
qo         : QOPeriods.find_periods default path, get_subspaces, solve, get_periods
ramanujan  : Cq, Cq_complete, float32-storing projection, find_periods_with_weights
"""
from . import numtheory, periods, qo, ramanujan  # noqa: F401
