"""Exact reproduction of ``numpy.random.RandomState.choice(n, p=<uniform>)`` without n-sized arrays.

scikit-learn draws the first k-means++ centre with ``random_state.choice(n, p=w / w.sum())``
(``sklearn/cluster/_kmeans.py:228``) and the ``init="random"`` seeds with
``random_state.choice(n, size=k, replace=False, p=w / w.sum())`` (``_kmeans.py:1014-1021``), with
unit weights ``w``.  numpy inverts one uniform draw ``u`` through

    cdf = np.cumsum(p); cdf /= cdf[-1]; index = cdf.searchsorted(u, side="right")

where ``p = np.full(n, 1.0 / n)``.  ``np.cumsum`` is a plain running sum, so ``cdf[i]`` is the result
of adding the double ``1.0 / n`` to zero ``i + 1`` times, one IEEE addition after the other.  For a
billion points that array does not fit the host (and would take seconds to build), but the
sequence has a closed form: while the running sum stays inside one binade its unit in the last
place is fixed, so every addition advances it by the SAME amount (after at most one step that
settles a round-half-to-even tie).  The sum is therefore piecewise linear in the index with a
few pieces per binade -- under a hundred between ``1/n`` and ``1`` -- found here with exact
rational arithmetic; a binary search over the index then reproduces ``searchsorted`` bit for bit.

Host-side arithmetic only (no device work): this is the random-number bookkeeping scikit-learn
does on the host as well.  Checked against numpy itself in ``tests/test_host.py``.
"""
from __future__ import annotations

from bisect import bisect_right
from fractions import Fraction
from typing import List, Tuple

import numpy as np


def _float_exact(x: Fraction) -> float:
    """A Fraction that is known to be a double, as that double."""
    f = float(x)
    assert Fraction(f) == x
    return f


class UniformCdf:
    """``np.cumsum(np.full(n, 1.0 / n))`` in closed form: ``at(i)`` is its element ``i``."""

    def __init__(self, n: int):
        if n < 1:
            raise ValueError("n must be >= 1")
        self.n = int(n)
        self.p = 1.0 / float(n)  # what np.full(n, 1.0 / n) holds
        # pieces: cdf[i] = s0 + (i - i0) * d  for i0 <= i < next piece's i0   (exact rationals)
        self._i0: List[int] = []
        self._piece: List[Tuple[int, Fraction, Fraction]] = []
        self._build()
        self.last = self.at(self.n - 1)

    def _build(self):
        p, n = self.p, self.n
        i, s = 0, p  # cdf[0] = 0.0 + p
        while True:
            self._add_piece(i, Fraction(s), Fraction(0))  # the element itself
            if i >= n - 1:
                return
            # Two plain IEEE steps first: a step that enters a binade, and the one after it, may
            # differ from the steady increment (change of ulp; round-half-to-even settling).
            s1 = s + p
            self._add_piece(i + 1, Fraction(s1), Fraction(0))
            if i + 1 >= n - 1:
                return
            s2 = s1 + p
            s3 = s2 + p
            d = Fraction(s3) - Fraction(s2)
            e2, e3 = np.frexp(s2)[1], np.frexp(s3)[1]
            if d == 0:
                # 1/n has dropped below half an ulp of the sum: it no longer moves
                self._add_piece(i + 2, Fraction(s2), Fraction(0))
                return
            if e3 != e2:
                i, s = i + 2, s2  # s3 is already in the next binade: keep stepping singly
                continue
            # From s2 on, inside its binade (fixed ulp), every addition advances the sum by d:
            # largest k with s2 + k * d strictly below the top of the binade
            top = Fraction(2) ** int(e2)
            k = -((Fraction(s2) - top) // d)  # ceil((top - s2) / d)
            k = int(k) - 1
            k = max(1, min(k, n - 1 - (i + 2)))
            self._add_piece(i + 2, Fraction(s2), d)
            i = i + 2 + k
            s = _float_exact(Fraction(s2) + k * d)

    def _add_piece(self, i0: int, s0: Fraction, d: Fraction):
        if self._i0 and self._i0[-1] == i0:
            self._piece[-1] = (i0, s0, d)
            return
        self._i0.append(i0)
        self._piece.append((i0, s0, d))

    def at(self, i: int) -> float:
        j = bisect_right(self._i0, i) - 1
        i0, s0, d = self._piece[j]
        return _float_exact(s0 + (i - i0) * d)

    def search(self, u: float, n_eff: int | None = None) -> int:
        """``(cdf[:n_eff] / cdf[n_eff - 1]).searchsorted(u, side="right")``: the first index whose
        normalised value exceeds ``u`` (``n_eff`` = how many leading elements take part)."""
        n_eff = self.n if n_eff is None else int(n_eff)
        last = np.float64(self.at(n_eff - 1))
        lo, hi = 0, n_eff  # invariant: value(lo - 1) <= u < value(hi)
        while lo < hi:
            mid = (lo + hi) // 2
            if np.float64(self.at(mid)) / last > u:
                hi = mid
            else:
                lo = mid + 1
        return lo


def choice_uniform(random_state, n: int, cdf: UniformCdf | None = None) -> int:
    """``random_state.choice(n, p=np.full(n, 1.0 / n))`` -- one draw from the stream, same result."""
    cdf = cdf or UniformCdf(n)
    u = random_state.random_sample()
    return int(min(cdf.search(u), n - 1))  # (numpy's searchsorted cannot return n here: cdf[-1] / cdf[-1] = 1 > u)


def choice_uniform_without_replacement(random_state, n: int, k: int, cdf: UniformCdf | None = None) -> np.ndarray:
    """``random_state.choice(n, size=k, replace=False, p=np.full(n, 1.0 / n))``.

    numpy (``mtrand.pyx``, the ``replace=False`` branch with ``p``) draws the missing indices in
    rounds: uniforms -> searchsorted on the cdf of ``p`` with the indices found so far zeroed ->
    first occurrences kept in draw order.  Zeroing an entry of ``p`` adds 0.0 at that index, which
    leaves the running sum unchanged: the cdf of the remaining entries is the SAME sequence with
    the zeroed indices skipped, so it is read from the closed form through an index map."""
    if k > n:
        raise ValueError("Cannot take a larger sample than population when 'replace=False'")
    cdf = cdf or UniformCdf(n)
    found = np.empty(0, dtype=np.int64)
    while found.size < k:
        x = random_state.random_sample(k - found.size)
        removed = np.sort(found)
        n_eff = n - removed.size
        new = np.empty(x.shape[0], dtype=np.int64)
        for t, u in enumerate(x):
            m = cdf.search(float(u), n_eff)  # index among the entries that are still there
            # the m-th surviving index: skip the removed ones at or below it
            j = m
            while True:
                jj = m + int(np.searchsorted(removed, j, side="right"))
                if jj == j:
                    break
                j = jj
            new[t] = j
        _, first = np.unique(new, return_index=True)
        first.sort()
        found = np.concatenate([found, new[first]])
    return found[:k]
