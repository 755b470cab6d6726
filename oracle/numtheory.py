"""Integer side of the oracle (test infrastructure; see oracle/__init__.py).

Follows the reference's helpers:
  get_factors  pyPeriod/Periods.py:55-84, pyPeriod/QOPeriods.py:46-75
  get_primes   pyPeriod/Periods.py:33-52 (PRIMES = primes <= 10000, Periods.py:121)
  phi          pyPeriod/QOPeriods.py:16-43
"""
from __future__ import annotations

from functools import lru_cache
from math import gcd

PRIME_LIMIT = 10000  # Periods.py:121


def divisor_set(n: int) -> set:
    """All divisors of n as a Python set built in the reference's insertion order.

    The reference iterates this set directly (Periods.py:209-210, 548-549), so the
    CPython hash-table iteration order is load-bearing for orthogonalisation and
    for M-best step 2.  The order depends on the insertion sequence, which is
    [1, n, 2, n//2, ...] for i = 1..int(n**0.5) (Periods.py:73-78).
    """
    n = int(n)
    seq = []
    for i in range(1, int(n ** 0.5) + 1):
        if n % i == 0:
            seq += [i, n // i]
    return set(seq)


def factors_in_set_order(n: int, remove_1: bool = True, remove_n: bool = True) -> list:
    """Divisors of n in the order the reference's `for f in get_factors(...)` sees them."""
    s = divisor_set(n)
    if remove_1:
        s.remove(1)  # KeyError for n == 1 with remove_n, as Periods.py:80-83
    if remove_n:
        s.remove(int(n))
    return list(s)


@lru_cache(maxsize=None)
def _prime_sieve(limit: int = PRIME_LIMIT) -> frozenset:
    flags = bytearray([1]) * (limit + 1)
    flags[0:2] = b"\x00\x00"
    for i in range(2, int(limit ** 0.5) + 1):
        if flags[i]:
            flags[i * i :: i] = bytearray(len(flags[i * i :: i]))
    return frozenset(i for i in range(limit + 1) if flags[i])


def is_listed_prime(f: int) -> bool:
    """Membership in the reference's PRIMES table (primes <= 10000 only)."""
    return int(f) in _prime_sieve()


def orth_chain(p: int) -> list:
    """Cofactors p//f for the prime divisors f of p, in the reference's loop order.

    Periods.py:208-214: `for f in get_factors(p) minus {1,p}: if f in PRIMES:
    projection -= project(projection, int(p / f), trunc, False)`.
    A prime p has no non-trivial divisors, hence an empty chain.
    """
    if p < 2:
        return []
    return [int(p / f) for f in factors_in_set_order(p) if is_listed_prime(f)]


@lru_cache(maxsize=None)
def phi(n: int) -> int:
    """Euler's totient by the reference's gcd count (QOPeriods.py:39-43)."""
    n = int(n)
    return sum(1 for k in range(1, n + 1) if gcd(n, k) == 1)
