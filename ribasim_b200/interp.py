"""Host-side interpolation objects used while compiling a model (init time only).

The reference delegates these to DataInterpolations.jl 9.2.0 (Manifest.toml:375-381), whose
source is not under /root/reference.  The arithmetic below restates its published behaviour:

* `ConstantInterpolation` (dir = :left): value `u[i]` on `[t_i, t_{i+1})`, Constant or
  Periodic extrapolation — call sites core/src/read.jl:84-148, core/src/util.jl:72-95.
* `LinearInterpolation` with cached slopes `(u[i+1]-u[i])/(t[i+1]-t[i])`,
  value `u[i] + slope*(t - t[i])`, cumulative integrals `I`.
* `invert_integral(LinearInterpolation)` -> `LinearInterpolationIntInv`
  (core/src/read.jl:2289-2296): `h = h_i + (-a_i + sqrt(a_i^2 + 2 m_i dS)) / m_i`
  (or `dS / a_i` when the slope is zero); this form reproduces the reference's golden
  level 0.04471158417652035 for S = 1 on A=[0.01,1000], h=[0,1] bit-for-bit.
* `PCHIPInterpolation` = `CubicHermiteSpline(du_PCHIP(u, t), u, t)` — call sites
  core/src/util.jl:134-140, core/src/read.jl:1188-1193.  The knot slopes follow the classic
  Fritsch-Carlson / Moler `pchip` rules; core/test/validation_test.jl:7-15 pins the end-slope
  rule (itp(1.5) = 0.03125, itp(3.0) = 0.25) and the `trivial` end state
  (core/regression_test/regression_test.jl:35-41) pins it again through a full run.
"""

from __future__ import annotations

import bisect
import math

import numpy as np

PREVFLOAT_INF = math.nextafter(math.inf, 0.0)


def searchsortedlast(a, x) -> int:
    """1-based index of the last element <= x (0 if none), like Julia's."""
    return bisect.bisect_right(a, x)


def searchsortedfirst(a, x) -> int:
    """1-based index of the first element >= x (len+1 if none), like Julia's."""
    return bisect.bisect_left(a, x) + 1


class TimeSeries:
    """Scalar time interpolation: kind 'constant' (forward fill) or 'linear'."""

    def __init__(self, u, t, kind: str = "constant", periodic: bool = False):
        self.u = [float(x) for x in u]
        self.t = [float(x) for x in t]
        assert len(self.u) == len(self.t) >= 1
        self.kind = kind
        self.periodic = periodic
        n = len(self.t)
        if kind == "linear":
            self.slope = [
                (self.u[i + 1] - self.u[i]) / (self.t[i + 1] - self.t[i]) for i in range(n - 1)
            ]
        # cumulative integral at the knots (cache_parameters = true)
        I = [0.0] * n
        for i in range(n - 1):
            I[i + 1] = I[i] + self._segment_integral(i, self.t[i], self.t[i + 1])
        self.I = I

    @classmethod
    def static(cls, val: float, kind: str = "constant") -> "TimeSeries":
        # core/src/read.jl:118-120: [val, val] on [0, prevfloat(Inf)]
        return cls([val, val], [0.0, PREVFLOAT_INF], kind=kind)

    # -- helpers -----------------------------------------------------------------
    def _wrap(self, t: float) -> float:
        if self.periodic and (t < self.t[0] or t > self.t[-1]):
            T = self.t[-1] - self.t[0]
            return (t - self.t[0]) % T + self.t[0]
        return t

    def _segment_integral(self, i: int, t1: float, t2: float) -> float:
        if self.kind == "constant":
            return self.u[i] * (t2 - t1)
        u_mean = self.u[i] + self.slope[i] * ((t1 + t2) / 2 - self.t[i])
        return u_mean * (t2 - t1)

    # -- evaluation --------------------------------------------------------------
    def __call__(self, t: float) -> float:
        t = self._wrap(t)
        n = len(self.t)
        if self.kind == "constant":
            if t < self.t[0]:
                return self.u[0]
            idx = min(max(searchsortedlast(self.t, t), 1), n)
            return self.u[idx - 1]
        if t < self.t[0]:
            return self.u[0]
        if t > self.t[-1]:
            return self.u[-1]
        idx = min(max(searchsortedlast(self.t, t), 1), n - 1)
        return self.u[idx - 1] + self.slope[idx - 1] * (t - self.t[idx - 1])

    def derivative(self, t: float) -> float:
        if self.kind == "constant":
            return 0.0
        t = self._wrap(t)
        if t < self.t[0] or t > self.t[-1]:
            return 0.0
        idx = min(max(searchsortedlast(self.t, t), 1), len(self.t) - 1)
        return self.slope[idx - 1]

    def integral(self, t1: float, t2: float) -> float:
        """Exact integral over [t1, t2] (core/src/solve.jl:95).  Constant extrapolation
        outside the data; periodic series are integrated piecewise per period."""
        if t1 == t2:
            return 0.0
        if t1 > t2:
            return -self.integral(t2, t1)
        if self.periodic:
            T = self.t[-1] - self.t[0]
            total = 0.0
            a = t1
            while a < t2:
                wa = (a - self.t[0]) % T + self.t[0]
                room = self.t[-1] - wa
                step = min(room, t2 - a)
                total += self._integral_inside(wa, wa + step)
                a += step
            return total
        total = 0.0
        lo, hi = self.t[0], self.t[-1]
        if t1 < lo:
            total += self.u[0] * (min(t2, lo) - t1)
            t1 = lo
            if t2 <= lo:
                return total
        if t2 > hi:
            total += self.u[-1] * (t2 - max(t1, hi))
            t2 = hi
            if t1 >= hi:
                return total
        return total + self._integral_inside(t1, t2)

    def _integral_inside(self, t1: float, t2: float) -> float:
        n = len(self.t)
        idx1 = min(max(searchsortedlast(self.t, t1), 1), n - 1)
        idx2 = min(max(searchsortedfirst(self.t, t2) - 1, 1), n - 1)
        if idx1 == idx2:
            return self._segment_integral(idx1 - 1, t1, t2)
        total = self._segment_integral(idx1 - 1, t1, self.t[idx1])
        total += self._segment_integral(idx2 - 1, self.t[idx2 - 1], t2)
        total += self.I[idx2 - 1] - self.I[idx1]
        return total

    def transition_ts(self) -> list[float]:
        return list(self.t)

    def tstops(self, t_end: float) -> list[float]:
        """core/src/util.jl:1098-1129."""
        ts = self.transition_ts()
        if not self.periodic:
            return [x for x in ts if 0 <= x <= t_end]
        T = ts[-1] - ts[0]
        n_back = int(math.ceil(ts[0] / T))
        n_fwd = int(math.ceil((t_end - ts[0]) / T))
        out: list[float] = []
        for i in range(-n_back, n_fwd + 1):
            src = ts if i == n_fwd else ts[:-1]
            out.extend(x + i * T for x in src if 0 <= x + i * T <= t_end)
        return out


class IndexLookup(TimeSeries):
    """Time -> integer index lookup (ConstantInterpolation over Int), read.jl:552-558."""

    def __call__(self, t: float) -> int:  # type: ignore[override]
        return int(super().__call__(t))


# ------------------------------------------------------------------------- profiles
def trapz_integrate(dfdx, x):
    """core/src/read.jl:2188-2197."""
    n = len(dfdx)
    f = [0.0] * n
    for i in range(n - 1):
        dx = x[i + 1] - x[i]
        f[i + 1] = f[i] + 0.5 * (dfdx[i + 1] + dfdx[i]) * dx
    return f


def finite_difference(f, x, dfdx0=None):
    """core/src/read.jl:2171-2185."""
    if dfdx0 is None:
        dfdx0 = (f[1] - f[0]) / (x[1] - x[0])
    dfdx = [0.0] * len(x)
    dfdx[0] = dfdx0
    for i in range(len(x) - 1):
        df = f[i + 1] - f[i]
        dx = x[i + 1] - x[i]
        dfdx[i + 1] = 2 * df / dx - dfdx[i]
    return dfdx


class LinearTable:
    """LinearInterpolation over a non-time axis with cached slopes and integrals.

    extrapolation: 'constant' | 'extension' per side (read.jl:2281-2287, 2305-2310).
    """

    def __init__(self, u, t, left: str = "constant", right: str = "constant"):
        self.u = [float(x) for x in u]
        self.t = [float(x) for x in t]
        n = len(self.t)
        self.slope = [
            (self.u[i + 1] - self.u[i]) / (self.t[i + 1] - self.t[i]) for i in range(n - 1)
        ]
        self.left, self.right = left, right
        I = [0.0] * n
        for i in range(n - 1):
            I[i + 1] = I[i] + self._seg_int(i, self.t[i], self.t[i + 1])
        self.I = I

    def _seg_int(self, i, t1, t2):
        u_mean = self.u[i] + self.slope[i] * ((t1 + t2) / 2 - self.t[i])
        return u_mean * (t2 - t1)

    def __call__(self, t: float) -> float:
        n = len(self.t)
        if t < self.t[0] and self.left == "constant":
            return self.u[0]
        if t > self.t[-1] and self.right == "constant":
            return self.u[-1]
        idx = min(max(searchsortedlast(self.t, t), 1), n - 1)
        return self.u[idx - 1] + self.slope[idx - 1] * (t - self.t[idx - 1])

    def integral_from_start(self, t: float) -> float:
        """`integral(A, t)` = integral from t[1] to t (used by get_storage_from_level,
        core/src/util.jl:2-9, for t >= t[1])."""
        t1 = self.t[0]
        if t == t1:
            return 0.0
        n = len(self.t)
        total = 0.0
        t2 = t
        if t2 > self.t[-1]:
            # right extrapolation part
            d = t2 - self.t[-1]
            if self.right == "constant":
                total += self.u[-1] * d
            else:  # extension of the last segment
                s = self.slope[-1]
                total += (self.u[-1] + 0.5 * s * d) * d
            t2 = self.t[-1]
            if n == 1:
                return total
            return total + self.I[-1]
        idx2 = min(max(searchsortedfirst(self.t, t2) - 1, 1), n - 1)
        if idx2 == 1:
            return total + self._seg_int(0, t1, t2)
        total += self._seg_int(0, t1, self.t[1])
        total += self._seg_int(idx2 - 1, self.t[idx2 - 1], t2)
        total += self.I[idx2 - 1] - self.I[1]
        return total


def storage_to_level(S: float, level, dSdh, slope, cumS) -> float:
    """LinearInterpolationIntInv with Linear extrapolation on both sides
    (core/src/read.jl:2289-2296, evaluated at core/src/util.jl:42-44)."""
    n = len(level)
    if S < cumS[0]:
        return level[0] + (S - cumS[0]) / dSdh[0]
    if S > cumS[-1]:
        return level[-1] + (S - cumS[-1]) / dSdh[-1]
    idx = min(max(searchsortedlast(cumS, S), 1), n - 1) - 1
    dS = S - cumS[idx]
    x = dSdh[idx]
    m = slope[idx]
    # cancellation-free quadratic root (finite-difference slopes of constant-area segments are
    # ~1e-15, not 0: core/src/read.jl:2171-2185)
    return level[idx] + 2 * dS / (x + math.sqrt(x * x + m * (2 * dS)))


# ------------------------------------------------------------------------- PCHIP
def _pchip_end(h1, h2, del1, del2):
    """Three-point end slope with the shape-preserving caps of Moler's `pchipend`."""
    d = ((2 * h1 + h2) * del1 - h1 * del2) / (h1 + h2)
    sgn = lambda x: int(x > 0) - int(x < 0)  # noqa: E731
    if sgn(d) != sgn(del1):
        return 0.0
    if sgn(del1) != sgn(del2) and abs(d) > abs(3 * del1):
        return 3 * del1
    return d


def du_pchip(u, t):
    """Knot slopes of PCHIPInterpolation (Fritsch-Carlson / Moler `pchip`).

    Pinned by core/test/validation_test.jl:7-15: level=[1,2], Q=[0,0.1] (after the prepended
    point) must give itp(1.5) = 0.03125 and itp(3.0) = 0.25, i.e. an end slope of 1.5*delta."""
    n = len(t)
    h = [t[i + 1] - t[i] for i in range(n - 1)]
    delta = [(u[i + 1] - u[i]) / h[i] for i in range(n - 1)]
    du = [0.0] * n
    if n == 2:
        return [delta[0], delta[0]]
    for k in range(1, n - 1):
        a, b = delta[k - 1], delta[k]
        if a == 0.0 or b == 0.0 or (a > 0) != (b > 0):
            du[k] = 0.0
        else:
            w1 = 2 * h[k] + h[k - 1]
            w2 = h[k] + 2 * h[k - 1]
            du[k] = a * b * (w1 + w2) / (w1 * b + w2 * a)
    du[0] = _pchip_end(h[0], h[1], delta[0], delta[1])
    du[-1] = _pchip_end(h[-1], h[-2], delta[-1], delta[-2])
    return du


class PchipTable:
    """CubicHermiteSpline with cached (c1, c2) per segment.

    left/right: 'constant' or 'linear' extrapolation (util.jl:134-140, read.jl:1188-1193).
    """

    def __init__(self, u, t, left: str, right: str):
        self.u = [float(x) for x in u]
        self.t = [float(x) for x in t]
        self.du = du_pchip(self.u, self.t)
        n = len(self.t)
        self.c1, self.c2 = [], []
        for i in range(n - 1):
            dt = self.t[i + 1] - self.t[i]
            c1 = (self.u[i + 1] - self.u[i] - self.du[i] * dt) / dt**2
            c2 = (self.du[i + 1] - self.du[i] - 2 * c1 * dt) / dt**2
            self.c1.append(c1)
            self.c2.append(c2)
        self.left, self.right = left, right
        # slopes used by Linear extrapolation = derivative at the end knots
        self.slope_left = self.du[0] if left == "linear" else 0.0
        if right == "linear":
            i = n - 2
            dt0 = self.t[-1] - self.t[i]
            self.slope_right = self.du[i] + dt0 * (dt0 * self.c2[i] + 2 * (self.c1[i] + 0.0 * self.c2[i]))
        else:
            self.slope_right = 0.0

    def __call__(self, x: float) -> float:
        n = len(self.t)
        if x < self.t[0]:
            return self.u[0] + self.slope_left * (x - self.t[0])
        if x > self.t[-1]:
            return self.u[-1] + self.slope_right * (x - self.t[-1])
        i = min(max(searchsortedlast(self.t, x), 1), n - 1) - 1
        d0 = x - self.t[i]
        d1 = x - self.t[i + 1]
        out = self.u[i] + d0 * self.du[i]
        out += d0**2 * (self.c1[i] + d1 * self.c2[i])
        return out


def as_f64(x) -> np.ndarray:
    return np.ascontiguousarray(np.asarray(x, dtype=np.float64))


def as_i32(x) -> np.ndarray:
    return np.ascontiguousarray(np.asarray(x, dtype=np.int32))
