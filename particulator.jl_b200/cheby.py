"""Chebyshev interpolation in binary intervals — host-side table construction.

Mirrors src/cheby.jl of the reference (BinaryIntervals :15-21, interval :24-36, chebnodes
:41-52, chebval :57-81, chebeval :111-125, precheb :127-143, chebdiff :180-203, chebfit
:211-227).  This is *init* code: it runs once on the host and hands flat coefficient arrays to
the device library; the per-particle evaluation lives in csrc/ (and, independently, in oracle/)."""
from dataclasses import dataclass
import math
import numpy as np


@dataclass(frozen=True)
class BinaryIntervals:
    """cheby.jl:15-21: intervals indexed 0..k, root interval (0, xmax)."""
    k: int
    xmax: float

    def interval(self, i):
        """cheby.jl:24-36"""
        if i == 0:
            l, _ = self.interval(1)
            return (0.0, l)
        l = 2.0 ** (i - self.k - 1)
        return (l * self.xmax, 2 * l * self.xmax)


def chebnodes(n, a=-1.0, b=1.0):
    """cheby.jl:41-43"""
    kk = np.arange(n)
    return (a + b) / 2 + (b - a) / 2 * np.cos((2 * kk + 1) * math.pi / (2 * n))


def chebval(xi, n):
    """cheby.jl:57-81: (T_0(xi), ..., T_{n-1}(xi)) by the three-term recurrence."""
    xi = np.asarray(xi, dtype=np.float64)
    t = [np.ones_like(xi)]
    if n > 1:
        t.append(xi)
    for _ in range(2, n):
        t.append(2 * xi * t[-1] - t[-2])
    return np.stack(t, axis=0)


def precheb(x, b: BinaryIntervals, n):
    """cheby.jl:127-143: interval index i and the Chebyshev values at the local coordinate."""
    x = np.asarray(x, dtype=np.float64)
    x1 = x / b.xmax
    s, l = np.frexp(x1)
    i = np.where(x1 == 0, 0, l + b.k)
    xi = np.where(i > 0, 4 * s - 3, 2.0 ** (b.k + 1) * x1 - 1)
    i = np.where(i > 0, i, 0)
    return i.astype(np.int64), chebval(xi, n)


def chebeval(x, b: BinaryIntervals, a):
    """cheby.jl:111-119: a[order, k+1]"""
    i, t = precheb(x, b, a.shape[0])
    return np.sum(a[:, i] * t, axis=0)


def chebdiff(x, b: BinaryIntervals, a):
    """cheby.jl:180-203: derivative of the expansion (used only for rate bounds)."""
    x = np.asarray(x, dtype=np.float64)
    x1 = x / b.xmax
    s, l = np.frexp(x1)
    i = np.where(x1 == 0, 0, l + b.k)
    xi = np.where(i > 0, 4 * s - 3, 2.0 ** (b.k + 1) * x1 - 1)
    i = np.where(i > 0, i, 0).astype(np.int64)
    lo = np.where(i == 0, 0.0, 2.0 ** (i - b.k - 1) * b.xmax)
    hi = np.where(i == 0, 2.0 ** (1 - b.k - 1) * b.xmax, 2 * 2.0 ** (i - b.k - 1) * b.xmax)
    u0 = np.ones_like(xi)
    u1 = 2 * xi
    df = a[1, i].copy()
    for j in range(3, a.shape[0] + 1):
        df = df + (j - 1) * u1 * a[j - 1, i]
        u0, u1 = u1, 2 * xi * u1 - u0
    return 2 * df / (hi - lo)


def chebfit(f, b: BinaryIntervals, n):
    """cheby.jl:211-227: collocation at the n Chebyshev nodes of every interval; a[n, k+1]."""
    a = np.zeros((n, b.k + 1))
    for i in range(b.k + 1):
        l, r = b.interval(i)
        x = chebnodes(n, l, r)
        with np.errstate(all="ignore"):
            fx = np.asarray(f(x), dtype=np.float64)
        xi = (2 * x - (l + r)) / (r - l)
        A = chebval(xi, n).T
        a[:, i] = np.linalg.solve(A, fx)
    return a
