"""Adaptive QNDF above the kernel seams.

The reference's default algorithm is `QNDF` (core/src/config.jl:368-380, `Solver.algorithm`): the
variable-order (1..5), variable-step numerical differentiation formulas of Shampine & Reichelt
("The MATLAB ODE suite", SIAM J. Sci. Comput. 18 (1997): `ode15s`), which OrdinaryDiffEqBDF's
`QNDF` implements (third-party, not under /root/reference; restated here from the paper):

    (1 - kappa_k) gamma_k (u_{n+1} - u0) + sum_{m=1..k} gamma_m D_m - h f(t_{n+1}, u_{n+1}) = 0,
    u0 = sum_{m=0..k} D_m,   D_m = m-th backward difference of the accepted solution,
    local error ~ (kappa_k gamma_k + 1/(k+1)) (u_{n+1} - u0),

with quasi-constant step sizes (the differences are re-interpolated when h changes) and the
order / step selection of `ode15s`.  The integrator owns the step control on the host and calls
the device through the SciML plug-in seams of the C ABI (include/ribasim_b200.h): `rbs_reduce`,
`rbs_rhs` (f!), `rbs_jac` (J_intermediate + W = A J - I/gamma, factorised), `rbs_solve` (the
reduced linear solve of `dolinsolve`, core/src/differentiation.jl:307-359), `rbs_limit_flow`
(`step_limiter!`) and `rbs_accept_step` (exact forcing bookkeeping).  The Newton system is solved
in the reduced space exactly like the reference: W c = A G / gamma, Delta = -G + gamma J_i c.

Tolerances follow the reference: weighted RMS norm with `abstol + max(|u_n|, |u_{n+1}|) reltol_i`,
a per-state `reltol` vector (core/src/util.jl:733-743) that `decrease_tolerance!`
(core/src/callback.jl:89-112) tightens as the cumulative states grow.  An ensemble advances in
lockstep with one step size and order: the norm is the maximum over the members, so every member
meets the tolerance (a single instance, the BMI case, is plain QNDF).
"""

from __future__ import annotations

import math

import numpy as np

MAXK = 5
KAPPA = np.array([-0.1850, -1.0 / 9.0, -0.0823, -0.0415, 0.0])
GAMMA = np.cumsum(1.0 / np.arange(1, MAXK + 1))
INV_GA = 1.0 / (GAMMA * (1.0 - KAPPA))
ERCONST = KAPPA * GAMMA + 1.0 / np.arange(2, MAXK + 2)
DIF_U = np.array([[-1, -2, -3, -4, -5],
                  [0, 1, 3, 6, 10],
                  [0, 0, -1, -4, -10],
                  [0, 0, 0, 1, 5],
                  [0, 0, 0, 0, -1]], dtype=float)
MAX_NEWTON = 4


def _dif_ru(ratio: float) -> np.ndarray:
    """Matrix that re-interpolates the backward differences from step h to ratio * h."""
    ki = np.arange(1, MAXK + 1)[:, None].astype(float)
    kj = np.arange(1, MAXK + 1)[None, :].astype(float)
    return np.cumprod((ki - 1.0 - kj * ratio) / ki, axis=0) @ DIF_U


class NewtonFailure(RuntimeError):
    pass


class Qndf:
    """One integrator for `n_members` instances in lockstep.  `backend` provides

        reduce(u) -> ur                       [n_reduced, B]
        rhs(ur, t) -> (du, storage)           [n, B], [n_basin, B]      (f! at time t)
        factor(ur, t, gamma) -> None          J_intermediate and the factors of W = A J - I/gamma
        solve(b_r) -> c                       W c = b_r
        jmul(c) -> J_intermediate c           [n, B]
        limit(u, uprev, dt, t) -> u           step_limiter!
        prepare(t0, t1)                       time-dependent parameters of the attempt [t0, t1]
    """

    def __init__(self, backend, u0: np.ndarray, t0: float, abstol: float, reltol: float,
                 relmask: np.ndarray, dtmin: float = 0.0, dtmax: float | None = None,
                 maxiters: int = 10**9, dt0: float | None = None):
        self.b = backend
        self.u = np.array(u0, dtype=float)
        self.n = self.u.shape[0]
        self.t = float(t0)
        self.abstol = float(abstol)
        self.reltol0 = float(reltol)
        self.reltol = np.full(self.n, float(reltol))
        self.relmask = np.asarray(relmask, dtype=bool)
        self.dtmin = float(dtmin)
        self.dtmax = math.inf if dtmax is None else float(dtmax)
        self.maxiters = int(maxiters)
        self.k = 1
        self.h = 0.0 if dt0 is None else float(dt0)
        self.dif = np.zeros(self.u.shape + (MAXK + 2,))
        self.nconhk = 0
        self.n_steps = self.n_rejected = self.n_rhs = self.n_factor = self.n_newton_fail = 0
        self._init_pending = True

    # ------------------------------------------------------------------ helpers
    def _scale(self, ua, ub):
        return self.abstol + np.maximum(np.abs(ua), np.abs(ub)) * self.reltol[:, None]

    def _norm(self, x, sc) -> float:
        r = x / sc
        return float(np.sqrt(np.mean(r * r, axis=0)).max())

    def _f(self, u, t):
        self.n_rhs += 1
        return self.b.rhs(self.b.reduce(u), t)

    def _rescale(self, h_new: float) -> None:
        if self.h > 0 and h_new != self.h:
            k = self.k
            self.dif[..., :k] = self.dif[..., :k] @ _dif_ru(h_new / self.h)[:k, :k]
        self.h = h_new
        self.nconhk = 0

    def _initialise(self, t_stop: float) -> None:
        """First step size (Hairer-Wanner's `hinit`, the recipe of OrdinaryDiffEq's initdt) and
        the first difference D_1 = h f(t0, u0)."""
        self.b.prepare(self.t, self.t)
        f0, _ = self._f(self.u, self.t)
        sc = self._scale(self.u, self.u)
        span = max(t_stop - self.t, 1e-12)
        if self.h <= 0.0:
            d0, d1 = self._norm(self.u, sc), self._norm(f0, sc)
            h0 = 1e-6 if (d0 < 1e-5 or d1 < 1e-5) else 0.01 * d0 / d1
            h0 = min(h0, span, self.dtmax)
            self.b.prepare(self.t, self.t + h0)
            f1, _ = self._f(self.u + h0 * f0, self.t + h0)
            d2 = self._norm(f1 - f0, sc) / h0
            h1 = max(1e-6, 1e-3 * h0) if max(d1, d2) <= 1e-15 else (0.01 / max(d1, d2)) ** 0.5
            self.h = min(100.0 * h0, h1, span, self.dtmax)
        self.h = max(self.h, self.dtmin, 1e-12)
        self.dif[..., 0] = self.h * f0
        self.k = 1
        self.nconhk = 0
        self._init_pending = False

    def restart(self) -> None:
        """After a discontinuity the caller cannot smooth over: order 1, new first step."""
        self.h = 0.0
        self._init_pending = True

    # ------------------------------------------------------------------ one accepted step
    def step(self, t_stop: float) -> float:
        """Advance by one accepted step that does not pass `t_stop`; returns its length."""
        if self.t >= t_stop:
            return 0.0
        if self._init_pending:
            self._initialise(t_stop)
        b = self.b
        n_fail = 0
        h_min = max(self.dtmin, 16 * np.finfo(float).eps * max(abs(self.t), 1.0))
        while True:
            if self.n_steps + self.n_rejected >= self.maxiters:
                raise RuntimeError("maxiters reached")
            # never step past the stop; a step that would leave a sliver takes the stop itself
            h_try = min(self.h, self.dtmax)
            rem = t_stop - self.t
            clipped_from = None
            if h_try >= rem * (1.0 - 1e-12) or rem - h_try < 1e-3 * h_try:
                clipped_from = h_try if h_try > rem else None
                h_try = rem
            if h_try != self.h:
                self._rescale(h_try)
            h, k = self.h, self.k
            t_new = t_stop if h == rem else self.t + h
            gamma = h * INV_GA[k - 1]
            b.prepare(self.t, t_new)
            # predictor
            psi = (self.dif[..., :k] * (GAMMA[:k] * INV_GA[k - 1])).sum(axis=-1)
            pred = self.u + self.dif[..., :k].sum(axis=-1)
            u_new = pred.copy()
            difkp1 = np.zeros_like(pred)
            sc = self._scale(self.u, u_new)
            # modified Newton on the reduced system, W factorised once per attempt at the predictor
            ur = b.reduce(u_new)
            b.factor(ur, t_new, gamma)
            self.n_factor += 1
            converged, rate, old_nrm = False, None, None
            min_nrm = 100 * np.finfo(float).eps * self._norm(u_new, sc)
            for it in range(1, MAX_NEWTON + 1):
                self.n_rhs += 1
                du, _ = b.rhs(ur, t_new)
                g = psi + difkp1 - gamma * du          # G(u) = (u - u0) + psi - gamma f(u)
                c = b.solve(b.reduce(g) / gamma)
                delta = -g + gamma * b.jmul(c)
                if not np.all(np.isfinite(delta)):
                    break
                nrm = self._norm(delta, sc)
                difkp1 = difkp1 + delta
                u_new = pred + difkp1
                ur = b.reduce(u_new)
                if nrm <= min_nrm:
                    converged = True
                elif it == 1:
                    if rate is not None and nrm * rate / (1 - rate) <= 0.05:
                        converged = True
                    else:
                        rate = None
                elif nrm > 0.9 * old_nrm:
                    break
                else:
                    rate = max(0.9 * rate, nrm / old_nrm) if rate is not None else nrm / old_nrm
                    errit = nrm * rate / (1 - rate)
                    if errit <= 0.05:
                        converged = True
                    elif it == MAX_NEWTON or 0.05 < errit * rate ** (MAX_NEWTON - it):
                        break
                if converged:
                    break
                old_nrm = nrm
            if not converged:
                self.n_newton_fail += 1
                self.n_rejected += 1
                if h <= h_min * (1 + 1e-12):
                    raise NewtonFailure(f"Newton iteration failed at t = {self.t} with the smallest step {h}")
                self._rescale(max(0.3 * h, h_min))
                n_fail += 1
                continue
            # step_limiter! on the new state, then the local error of order k
            u_lim = b.limit(u_new, self.u, h, t_new)
            if u_lim is not u_new:
                difkp1 = difkp1 + (u_lim - u_new)
                u_new = u_lim
            sc = self._scale(self.u, u_new)
            err = self._norm(difkp1, sc) * abs(ERCONST[k - 1])
            _, storage = b.rhs(b.reduce(u_new), t_new)
            self.n_rhs += 1
            out_of_domain = bool(np.any(storage < 0.0))
            if err > 1.0 or out_of_domain:
                self.n_rejected += 1
                if h <= h_min * (1 + 1e-12):
                    if out_of_domain:
                        raise RuntimeError(f"negative storage at t = {t_new} with the smallest step")
                    raise RuntimeError(f"step size underflow at t = {self.t}")
                if out_of_domain and err <= 1.0:
                    h_opt = 0.5 * h
                elif n_fail == 0:
                    h_opt = h * max(0.1, 0.833 * (1.0 / err) ** (1.0 / (k + 1)))
                    if k > 1:
                        err_km1 = self._norm(self.dif[..., k - 1] + difkp1, sc) * abs(ERCONST[k - 2])
                        h_km1 = h * max(0.1, 0.769 * (1.0 / err_km1) ** (1.0 / k))
                        if h_km1 > h_opt:
                            h_opt = min(h, h_km1)
                            self.k = k - 1
                else:
                    h_opt = 0.5 * h
                self._rescale(max(h_opt, h_min))
                n_fail += 1
                continue
            break
        # ---- accepted: new backward differences
        self.n_steps += 1
        k = self.k
        self.dif[..., k + 1] = difkp1 - self.dif[..., k]
        self.dif[..., k] = difkp1
        for j in range(k - 1, -1, -1):
            self.dif[..., j] += self.dif[..., j + 1]
        self.nconhk = min(self.nconhk + 1, MAXK + 2)
        h_done = h
        # ---- order and step size of the next step
        if self.nconhk >= k + 2:
            temp = 1.2 * err ** (1.0 / (k + 1))
            h_opt = h / temp if temp > 0.1 else 10.0 * h
            k_opt = k
            if k > 1:
                err_km1 = self._norm(self.dif[..., k - 1], sc) * abs(ERCONST[k - 2])
                temp = 1.3 * err_km1 ** (1.0 / k)
                h_km1 = h / temp if temp > 0.1 else 10.0 * h
                if h_km1 > h_opt:
                    h_opt, k_opt = h_km1, k - 1
            if k < MAXK:
                err_kp1 = self._norm(self.dif[..., k + 1], sc) * abs(ERCONST[k])
                temp = 1.4 * err_kp1 ** (1.0 / (k + 2))
                h_kp1 = h / temp if temp > 0.1 else 10.0 * h
                if h_kp1 > h_opt:
                    h_opt, k_opt = h_kp1, k + 1
            if h_opt > h and n_fail == 0:
                self.k = k_opt
                self._rescale(min(h_opt, self.dtmax))
        if clipped_from is not None and n_fail == 0 and self.h < clipped_from:
            self._rescale(min(clipped_from, self.dtmax))  # the stop cut the step short, not the error
        self.u = u_new
        self.t = t_new
        return h_done

    # ------------------------------------------------------------------ reference callbacks
    def decrease_tolerance(self) -> None:
        """`decrease_tolerance!` (core/src/callback.jl:89-112): the cumulative states grow without
        bound, so their relative tolerance shrinks with their magnitude over the mean magnitude."""
        if self.t <= 0.0:
            return
        mag = np.abs(self.u).max(axis=1)
        avg = np.maximum(1.0e4, mag / self.t)
        with np.errstate(divide="ignore", invalid="ignore"):
            diff = np.where(mag > 0.0, np.maximum(0.0, np.log10(mag / avg)), 0.0)
        new = np.maximum(10.0 ** (math.log10(self.reltol0) - diff), 1.0e-14)
        upd = self.relmask & (mag > 0.0) & (self.reltol > new)
        self.reltol[upd] = new[upd]
