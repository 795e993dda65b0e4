"""Adaptive QNDF (ribasim_b200/qndf.py) on the CPU oracle as backend: the integrator's logic is
checked here without a GPU against the reference's own regression values for QNDF
(core/regression_test/regression_test.jl:8-41, core/test/run_models_test.jl:221-222); the same
integrator on the device seams is checked in tests/test_configs_gpu.py."""

from __future__ import annotations

import numpy as np
import pytest

from oracle.oracle import OracleModel
from ribasim_b200.flatten import flatten
from ribasim_b200.network import build_params
from ribasim_b200.qndf import ERCONST, GAMMA, KAPPA, Qndf, _dif_ru
from ribasim_b200.testmodels import MODELS
from ribasim_b200.timecache import (eval_time_params, flow_boundary_increment, forcing_tstops,
                                    update_vertical_flux)
from tests.helpers import golden

FORCING = ("precipitation", "surface_runoff", "potential_evaporation", "drainage", "infiltration")


class OracleBackend:
    """The seams of include/ribasim_b200.h on the CPU restatement (one instance)."""

    def __init__(self, p, flat):
        self.p, self.flat = p, flat
        self.om = OracleModel(p)
        self.om.set_pattern(flat.jac_rows_csc, flat.jac_cols_csc)
        n = flat["n_state"]
        eye = np.eye(n)
        self.A = np.stack([self.om.reduce(eye[:, s]) for s in range(n)], axis=1)
        self.tprev = 0.0
        self.push_forcing()

    def push_forcing(self):
        for k in FORCING:
            self.om.set_f64(k, self.p.basin.vertical_flux[k])

    def prepare(self, t0, t1):
        self.om.set_time_params(eval_time_params(self.p, t1))
        self.om.set_f64("fb_incr", flow_boundary_increment(self.p, self.tprev, t1))

    def reduce(self, u):
        return self.om.reduce(np.ascontiguousarray(u[:, 0]))[:, None]

    def rhs(self, ur, t):
        du, c = self.om.rhs(np.ascontiguousarray(ur[:, 0]), t, with_cache=True)
        return du[:, None], c["storage"][:, None]

    def factor(self, ur, t, gamma):
        self.J = self.om.jac_dense(np.ascontiguousarray(ur[:, 0]), t)
        self.W = self.A @ self.J - np.eye(self.A.shape[0]) / gamma

    def solve(self, b):
        return np.linalg.solve(self.W, b)

    def jmul(self, c):
        return self.J @ c

    def limit(self, u, uprev, dt, t):
        return self.om.limit_flow(np.ascontiguousarray(u[:, 0]), np.ascontiguousarray(uprev[:, 0]), dt, t)[:, None]

    def accept(self, dt, t_new):
        """update_cumulative_flows! (core/src/callback.jl:131-170)."""
        om, nb, nfb = self.om, self.flat["n_basin"], self.flat["n_fb"]
        vf = self.p.basin.vertical_flux
        fixed_area = om.get_f64("fixed_area", nb)
        om.set_f64("cumulative_precipitation", om.get_f64("cumulative_precipitation", nb) + fixed_area * vf["precipitation"] * dt)
        om.set_f64("cumulative_surface_runoff", om.get_f64("cumulative_surface_runoff", nb) + dt * vf["surface_runoff"])
        om.set_f64("cumulative_drainage", om.get_f64("cumulative_drainage", nb) + dt * vf["drainage"])
        if nfb:
            om.set_f64("fb_cumulative_flow", om.get_f64("fb_cumulative_flow", nfb)
                       + flow_boundary_increment(self.p, self.tprev, t_new))
            om.set_f64("fb_incr", np.zeros(nfb))
        self.tprev = t_new
        om.L.orc_set_tprev(om.h, t_new)


def run_qndf(name, abstol=None, reltol=None, t_end=None):
    p = build_params(MODELS[name]())
    flat = flatten(p)
    m = p.tables
    t_end = m.t_end if t_end is None else t_end
    be = OracleBackend(p, flat)
    n = flat["n_state"]
    relmask = np.ones(n, dtype=bool)
    relmask[flat["off_integral"]:] = False
    q = Qndf(be, np.zeros((n, 1)), 0.0, abstol or m.solver_option("abstol"), reltol or m.solver_option("reltol"), relmask)
    stops = [t for t in sorted(set(p.tstops) | set(forcing_tstops(p, t_end))) if 0.0 < t < t_end] + [t_end]
    fstops = set(forcing_tstops(p, t_end))
    for stop in stops:
        while q.t < stop:
            h = q.step(stop)
            be.accept(h, q.t)
            q.decrease_tolerance()
        if stop in fstops and update_vertical_flux(p, stop):
            be.push_forcing()
    be.prepare(q.t, q.t)
    _, S = be.rhs(be.reduce(q.u), q.t)
    return q, S[:, 0], be


def test_ndf_constants():
    # Shampine & Reichelt, table 1 (kappa) and the identities the formulas rest on
    assert np.allclose(KAPPA, [-0.1850, -1 / 9, -0.0823, -0.0415, 0.0])
    assert np.allclose(GAMMA, [1, 3 / 2, 11 / 6, 25 / 12, 137 / 60])
    assert np.allclose(ERCONST, KAPPA * GAMMA + 1 / np.arange(2, 7))
    # re-interpolating a polynomial's backward differences is exact: u(t) = t^2 sampled at step h
    h, r = 0.7, 0.35
    f = lambda t: t * t  # noqa: E731
    d = lambda step: np.array([f(0) - f(-step), f(0) - 2 * f(-step) + f(-2 * step)])  # noqa: E731
    assert np.allclose(d(h) @ _dif_ru(r)[:2, :2], d(r * h), atol=1e-14)


def test_trivial_regression_values_with_qndf():
    # core/regression_test/regression_test.jl:15-41 (abstol = reltol = 1e-7)
    q, S, be = run_qndf("trivial", abstol=1e-7, reltol=1e-7)
    g = golden("trivial_regression")
    assert S[0] == pytest.approx(g["storage"], abs=g["storage_atol"])
    level = be.om.L.orc_storage_to_level(be.om.h, 0, float(S[0]))
    assert level == pytest.approx(g["level"], abs=g["level_atol"])
    assert q.n_steps < 2000 and q.k >= 1


def test_basic_end_state_with_default_qndf():
    # core/test/run_models_test.jl:221-222: the reference's end storages come from its default
    # algorithm, QNDF with abstol = reltol = 1e-5
    q, S, _ = run_qndf("basic")
    g = golden("basic_end_storage")
    assert np.allclose(S, g["storage"], atol=g["atol"], rtol=0)
    # a year in a few hundred adaptive steps (the fixed-step path takes 8784)
    assert q.n_steps < 3000
