"""Pins the CPU oracle (oracle/rbs_oracle.cpp) to the reference's own known-answer tests.

No GPU.  Citations are relative to /root/reference.  The oracle is test infrastructure; these
tests are what allows it to stand in for the Julia core (which cannot run in this image).
"""

from __future__ import annotations

import math

import numpy as np
import pytest

from oracle.oracle import OracleModel, lib, run_oracle
from ribasim_b200.flatten import flatten
from ribasim_b200.interp import PchipTable, storage_to_level
from ribasim_b200.network import build_params
from ribasim_b200.tables import ModelTables
from ribasim_b200.testmodels import MODELS
from ribasim_b200.timecache import eval_time_params
from tests.helpers import golden


def test_reduction_factor_goldens():
    # core/test/utils_test.jl:331-342
    f = lib().orc_reduction_factor
    assert f(-2.0, 2.0) == 0.0
    assert f(0.0, 2.0) == 0.0
    assert f(1.0, 2.0) == 0.5
    assert f(3.0, 2.0) == 1.0
    assert f(math.inf, 2.0) == 1.0
    assert f(-math.inf, 2.0) == 0.0


def test_relaxed_root_goldens():
    # core/test/utils_test.jl:483-514
    r = lib().orc_relaxed_root
    assert r(0.0, 1.0e-3) == 0.0
    assert r(2.0, 1.0) == pytest.approx(math.sqrt(2.0), rel=1e-15)
    assert r(-2.0, 1.0) == pytest.approx(-math.sqrt(2.0), rel=1e-15)
    eps = 1.0e-3
    assert abs(r(eps, eps) - math.sqrt(eps)) < 1e-12
    assert abs(r(-eps, eps) + math.sqrt(eps)) < 1e-12
    assert abs(r(eps / 2, eps) - math.sqrt(eps / 2)) < 0.2
    assert r(1.0e-8, 1.0e-8) == pytest.approx(math.sqrt(1.0e-8), rel=1e-12)
    assert r(-1.0e-8, 1.0e-8) == pytest.approx(-math.sqrt(1.0e-8), rel=1e-12)


def _one_basin(area, level) -> ModelTables:
    m = MODELS["bucket"]()
    m.tables["Basin / profile"] = [dict(node_id=1, area=a, level=h) for a, h in zip(area, level)]
    m.tables["Basin / state"] = [dict(node_id=1, level=level[0] + 1e-3)]
    return m


def test_profile_lookup_goldens():
    # core/test/utils_test.jl:38-122: storage -> (area, level) on A=[1e-9,100,100], h=[0,10,15]
    p = build_params(_one_basin([1.0e-9, 100.0, 100.0], [0.0, 10.0, 15.0]))
    om = OracleModel(p)
    L = om.L
    lookup = lambda S: (L.orc_level_to_area(om.h, 0, L.orc_storage_to_level(om.h, 0, max(S, 0.0))),  # noqa: E731
                        L.orc_storage_to_level(om.h, 0, max(S, 0.0)))
    cumS = p.basin.cumS[0]
    assert cumS[0] == 0.0 and cumS[1] == pytest.approx(500.0, rel=1e-9)
    for S, A, h in zip(cumS, [1.0e-9, 100.0, 100.0], [0.0, 10.0, 15.0]):
        a, lv = lookup(S)
        assert lv == pytest.approx(h, abs=1e-12) and a == pytest.approx(A, rel=1e-9, abs=1e-12)
    assert lookup(-1.0) == (1.0e-9, 0.0)
    a, lv = lookup(100.0)
    assert lv == pytest.approx(math.sqrt(100.0 / 5), rel=1e-8)
    assert a == pytest.approx(10 * lv, rel=1e-8)
    a, lv = lookup(600.0)
    assert lv == pytest.approx(11.0, rel=1e-9) and a == 100.0
    # extrapolation above the table: constant area (core/src/read.jl:2303-2315)
    a, lv = lookup(1100.0)
    assert lv == pytest.approx(16.0, rel=1e-9) and a == 100.0


def test_profile_roundtrip():
    # core/test/utils_test.jl:124-189: level(storage) inverts storage(level)
    level = [0.0, 0.42601923740838954, 1.1726055542568279, 1.9918063978301288, 2.945965660308591,
             3.7918607426596513, 4.378609443214641, 4.500422081139986, 4.638188322915925,
             5.462975756944211]
    area = [0.5284895347829252, 0.7036603783547138, 0.6831597656207129, 0.7582032614294112,
            0.5718206017422349, 0.5390282084391234, 0.9650081130058792, 0.07071025361013983,
            0.10659325339342585, 1.1]
    p = build_params(_one_basin(area, level))
    om = OracleModel(p)
    b = p.basin
    cumS, lv_t, dS = np.array(b.cumS[0]), np.array(b.levels[0]), np.array(b.dSdh[0])
    for S in np.linspace(0.0, 2 * cumS[-1], 50):
        h = om.L.orc_storage_to_level(om.h, 0, float(S))
        assert h == storage_to_level(float(S), b.levels[0], b.dSdh[0], b.dSdh_slope[0], b.cumS[0])
        # storage from level: integral of the piecewise-linear area up to h
        if h >= lv_t[-1]:
            S_back = cumS[-1] + (h - lv_t[-1]) * dS[-1]
        else:
            i = max(int(np.searchsorted(lv_t, h, side="right")) - 1, 0)
            slope = (dS[i + 1] - dS[i]) / (lv_t[i + 1] - lv_t[i])
            S_back = cumS[i] + dS[i] * (h - lv_t[i]) + 0.5 * slope * (h - lv_t[i]) ** 2
        assert S_back == pytest.approx(S, rel=1e-10, abs=1e-10)


def _profiled(profiles: dict[int, dict]) -> ModelTables:
    """Unconnected Basins with the given profile columns (area and / or storage per level)."""
    m = ModelTables(starttime=MODELS["bucket"]().starttime, endtime=MODELS["bucket"]().endtime)
    for node_id, cols in profiles.items():
        m.add_node("Basin", node_id, profile=cols, state=dict(level=cols["level"][0] + 1e-3))
    return m


def test_basin_profile_initialisation_goldens():
    # core/test/read_test.jl:93-206 (`interpolate_basin_profile!` with area, storage or both)
    levels = [0.0, 1.0, 2.0, 3.0, 4.0, 5.0]
    areas = [(h + 1) * math.pi for h in levels]
    storages = [math.pi / 2 * ((h + 1) ** 2 - 1) for h in levels]
    p = build_params(_profiled({1: dict(level=levels, area=areas),
                                2: dict(level=levels, storage=storages),
                                3: dict(level=levels, area=areas, storage=storages)}))
    om = OracleModel(p)
    s2l = lambda i, S: om.L.orc_storage_to_level(om.h, i, S)  # noqa: E731
    assert s2l(0, storages[1]) == pytest.approx(s2l(2, storages[1]), rel=1e-8)
    assert s2l(0, storages[1]) == pytest.approx(s2l(1, storages[1]), rel=1e-8)
    assert om.L.orc_level_to_area(om.h, 0, levels[0]) == pytest.approx(om.L.orc_level_to_area(om.h, 2, levels[0]))
    # cylinder: S = 2000 above a 1000 m2 profile given up to 1 m -> level 2 (linear extrapolation)
    p = build_params(_profiled({1: dict(level=[0.0, 1.0], area=[1000.0, 1000.0])}))
    om = OracleModel(p)
    assert om.L.orc_storage_to_level(om.h, 0, 2000.0) == pytest.approx(2.0, rel=1e-8)
    # linear area
    p = build_params(_profiled({1: dict(level=[0.0, 1.0], area=[0.001, 1000.0])}))
    om = OracleModel(p)
    assert om.L.orc_storage_to_level(om.h, 0, 500.0005 + 1000.0) == pytest.approx(2.0, rel=1e-8)


def test_decreasing_area_profile_is_rejected():
    # core/test/read_test.jl:208-229
    cols = dict(level=[265.0, 270.0, 275.0, 280.0, 285.0, 287.0],
                storage=[0.0, 3.551e6, 1.6238e7, 4.5444e7, 1.06217e8, 1.08e8])
    with pytest.raises(ValueError, match=r"Invalid profile for Basin #1. The step from \(h=285.0, S=1.06217e8\) "
                                         r"to \(h=287.0, S=1.08e8\) implies a decreasing area"):
        build_params(_profiled({1: cols}))


def test_qh_interpolation_goldens():
    # core/test/validation_test.jl:1-44: the only values of the reference's PCHIP rating curve that
    # its tests pin (constant below the first level, cubic in between, linear above the last)
    from ribasim_b200.network import NodeID, qh_interpolation
    itp = qh_interpolation(NodeID("TabulatedRatingCurve", 1, 1), [1.0, 2.0], [0.0, 0.1])
    expect = {0.0: 0.0, 1.0: 0.0, 1.5: 0.03125, 2.0: 0.1, 3.0: 0.25}
    for x, q in expect.items():
        assert itp(x) == pytest.approx(q, rel=1e-12, abs=1e-300)
    # the same table through the oracle's evaluator (what the device is compared with)
    m = MODELS["rating_curve"]()
    for r, (h, q) in zip(m.tables["TabulatedRatingCurve / static"][:2], ((1.0, 0.0), (2.0, 0.1))):
        r["level"], r["flow_rate"] = h, q
    m.tables["TabulatedRatingCurve / static"] = m.tables["TabulatedRatingCurve / static"][:2]
    om = OracleModel(build_params(m))
    for x, q in expect.items():
        assert om.L.orc_pchip_eval(om.h, 0, x) == pytest.approx(q, rel=1e-12, abs=1e-300)
    # validation.jl (`valid_tabulated_curve_level` and friends): the reference's messages
    for lv, fl, msg in (([1.0, 2.0], [0.1, 0.2], "The `flow_rate` must start at 0."),
                        ([1.0, 1.0], [0.0, 0.1], "The `level` cannot be repeated."),
                        ([1.0, 2.0, 3.0], [0.0, 0.2, 0.1],
                         "The `flow_rate` cannot decrease with increasing `level`.")):
        with pytest.raises(ValueError, match=msg.replace("`", ".")):
            qh_interpolation(NodeID("TabulatedRatingCurve", 2, 1), lv, fl)


def test_pchip_pinned_against_scipy():
    """Interior slopes and values of the PCHIP tables, pinned against an independent implementation
    of the same published algorithm (Fritsch-Butland harmonic-mean interior slopes, three-point
    shape-preserving end slopes = DataInterpolations' `PCHIPInterpolation`,
    scipy.interpolate.PchipInterpolator): the host's `interp.py`, and the oracle's evaluator on a
    rating curve (with the prepended point (h0 - 1, 0), constant extrapolation below and linear
    above, core/src/util.jl:101-141).  Non-uniform tables of 2..11 points: monotone, with flat
    pieces, and arbitrary."""
    from scipy.interpolate import PchipInterpolator
    from ribasim_b200.interp import PchipTable, du_pchip

    rng = np.random.default_rng(3)
    for trial in range(200):
        n = int(rng.integers(2, 12))
        t = np.cumsum(rng.uniform(0.05, 3.0, size=n))
        kind = trial % 3
        if kind == 0:
            u = np.cumsum(rng.uniform(0, 2, size=n))
        elif kind == 1:
            u = rng.normal(size=n)
        else:
            u = np.concatenate([[0.0], np.cumsum(rng.choice([0.0, 1.0, 0.3], size=n - 1))])
        sp = PchipInterpolator(t, u)
        ds = sp.derivative()(t)
        d = np.array(du_pchip(list(u), list(t)))
        assert np.max(np.abs(d - ds) / np.maximum(1.0, np.abs(ds))) <= 1e-13, (trial, d, ds)
        tab = PchipTable(u, t, "constant", "linear")
        for x in rng.uniform(t[0], t[-1], size=40):
            assert tab(float(x)) == pytest.approx(float(sp(x)), rel=1e-13, abs=1e-13)
        # extrapolation: constant below, linear with the end-knot derivative above
        assert tab(t[0] - 1.7) == u[0]
        assert tab(t[-1] + 2.5) == pytest.approx(u[-1] + 2.5 * ds[-1], rel=1e-13, abs=1e-13)
    # the oracle's evaluator (what the device is compared with to 1e-12) on rating curves
    for seed in range(6):
        rng = np.random.default_rng(100 + seed)
        k = int(rng.integers(5, 9))
        level = 1.0 + np.concatenate([[0.0], np.cumsum(rng.uniform(0.1, 2.0, size=k - 1))])
        flow = np.concatenate([[0.0], np.cumsum(rng.uniform(0.0, 3.0, size=k - 1) * rng.choice([0.0, 1.0, 1.0], size=k - 1))])
        m = MODELS["rating_curve"]()
        proto = m.tables["TabulatedRatingCurve / static"][0]
        m.tables["TabulatedRatingCurve / static"] = [dict(proto, level=float(h), flow_rate=float(q))
                                                     for h, q in zip(level, flow)]
        om = OracleModel(build_params(m))
        ta, ua = np.concatenate([[level[0] - 1.0], level]), np.concatenate([[0.0], flow])
        sp = PchipInterpolator(ta, ua)
        for x in np.concatenate([rng.uniform(ta[0], ta[-1], size=60), level]):
            assert om.L.orc_pchip_eval(om.h, 0, float(x)) == pytest.approx(float(sp(x)), rel=1e-13, abs=1e-14)
        assert om.L.orc_pchip_eval(om.h, 0, float(ta[0] - 3.0)) == 0.0
        slope = float(sp.derivative()(ta[-1]))
        assert om.L.orc_pchip_eval(om.h, 0, float(ta[-1] + 1.5)) == pytest.approx(ua[-1] + 1.5 * slope, rel=1e-13, abs=1e-14)


def test_initial_level_golden():
    # python/ribasim_api/tests/test_bmi.py:102-110 and basic.py:45:
    # storage 1.0 on A=[0.01,1000], h=[0,1] <-> level 0.04471158417652035
    p = build_params(MODELS["basic"]())
    om = OracleModel(p)
    assert p.basin.storage0[0] == pytest.approx(1.0, rel=1e-12)
    h = om.L.orc_storage_to_level(om.h, 0, 1.0)
    assert h == pytest.approx(0.04471158417652035, rel=2e-15)
    # regression_test.jl:37-38: S = 62.2290641115 -> level 0.352778
    assert om.L.orc_storage_to_level(om.h, 0, 62.2290641115) == pytest.approx(0.352778, abs=1e-6)


def test_pchip_is_monotone_and_interpolates():
    """PCHIP (Fritsch-Carlson) properties DataInterpolations guarantees: hits the knots and
    preserves monotonicity; rating curves get the point (h0 - 1, 0) prepended
    (core/src/util.jl:130-140)."""
    t = [0.0, 1.0, 2.5, 3.0, 7.0]
    u = [0.0, 0.1, 2.0, 2.1, 30.0]
    tab = PchipTable(u, t, left="constant", right="linear")
    xs = np.linspace(0.0, 7.0, 2001)
    ys = np.array([tab(float(x)) for x in xs])
    assert np.all(np.diff(ys) >= -1e-14)
    for ti, ui in zip(t, u):
        assert tab(ti) == pytest.approx(ui, abs=1e-14)
    p = build_params(MODELS["rating_curve"]())
    trc = p.nodes["tabulated_rating_curve"]["interpolations"][0]
    assert trc.u[0] == 0.0 and trc.t[0] == pytest.approx(trc.t[1] - 1.0)
    om = OracleModel(p)
    for x in np.linspace(trc.t[0] - 0.5, trc.t[-1] + 1.0, 57):
        assert om.L.orc_pchip_eval(om.h, 0, float(x)) == pytest.approx(trc(float(x)), rel=1e-14,
                                                                       abs=1e-300)


def test_bucket_and_leaky_bucket_rhs():
    # core/test/run_models_test.jl:127-176
    p = build_params(MODELS["bucket"]())
    om = OracleModel(p)
    om.set_time_params(eval_time_params(p, 0.0))
    du, cache = om.rhs(np.zeros(om.n_reduced), 0.0, with_cache=True)
    assert cache["storage"][0] == pytest.approx(1000.0, rel=1e-12)
    assert du[flatten(p)["off_evap"]] == 0.0 and du[flatten(p)["off_infil"]] == 0.0


def test_trivial_end_state_implicit_euler():
    # core/regression_test/regression_test.jl:8,35-41 (ImplicitEuler is in the solver list)
    p = build_params(MODELS["trivial"]())
    u, S, lv, st = run_oracle(p, 86400.0 / 4, abstol=1e-7, reltol=1e-7)
    assert st["failed"] == 0
    g = golden("trivial_regression")
    assert S[0] == pytest.approx(g["storage"], abs=g["storage_atol"])
    assert lv[0] == pytest.approx(g["level"], abs=g["level_atol"])


def test_basic_end_state():
    # core/test/run_models_test.jl:221-222
    p = build_params(MODELS["basic"]())
    u, S, lv, st = run_oracle(p, 3600.0)
    g = golden("basic_end_storage")
    assert np.allclose(S, g["storage"], atol=g["atol"], rtol=0)
    assert st["negative"] == 0


def test_basic_transient_end_state():
    # core/test/run_models_test.jl:288-289
    p = build_params(MODELS["basic_transient"]())
    u, S, lv, st = run_oracle(p, 3600.0)
    g = golden("basic_transient_end_storage")
    assert np.allclose(S, g["storage"], atol=g["atol"], rtol=0)


def test_linear_resistance_analytic():
    # core/test/equations_test.jl:18-44: linear drain at Q_max, then exponential decay
    p = build_params(MODELS["linear_resistance"]())
    lr = p.nodes["linear_resistance"]
    R, Q_max = float(lr["resistance"][0]), float(lr["max_flow_rate"][0])
    A = p.basin.areas[0][-1]
    L = p.nodes["level_boundary"]["level"][0](0.0)
    u0 = A * 10.0
    t_shift = (u0 - A * (L + R * Q_max)) / Q_max
    for t_end in (30 * 86400.0, 366 * 86400.0):
        u, S, lv, st = run_oracle(p, 3600.0, t_end=t_end)
        if t_end < t_shift:
            expect = u0 - Q_max * t_end
        else:
            expect = A * L + A * R * Q_max * math.exp(-(t_end - t_shift) / (A * R))
        assert S[0] == pytest.approx(expect, rel=1e-4)


def _rating_curve_storage(t, S0, area):
    # core/test/equations_test.jl:46-70: w' = -alpha/A w^2
    storage_min, alpha = 50.005, 24 * 60 * 60
    return storage_min + 1.0 / (t / (alpha * area ** 2) + 1.0 / (S0 - storage_min))


def test_tabulated_rating_curve_analytic():
    # core/test/equations_test.jl:49-70 (rtol 0.01 on the whole trajectory; here several end times)
    p = build_params(MODELS["rating_curve"]())
    area = p.basin.areas[0][1]
    S0 = float(p.basin.storage0[0])
    for days in (1, 10, 100, 366):
        u, S, lv, st = run_oracle(p, 3600.0, t_end=days * 86400.0)
        assert S[0] == pytest.approx(_rating_curve_storage(days * 86400.0, S0, area), rel=0.01)


def _pid_equation_storage(p, t):
    # core/test/equations_test.jl:115-154
    pid = p.nodes["pid_control"]
    SP, K_p, K_i, K_d = (pid[k][0](0.0) for k in ("target", "proportional", "integral", "derivative"))
    storage_min, level_min = 50.005, p.basin.levels[0][1]
    S0, area = float(p.basin.storage0[0]), p.basin.areas[0][1]
    level0 = level_min + (S0 - storage_min) / area
    a, b, g = 1 - K_d / area, -K_p / area, -K_i / area
    d = -K_i * (SP - level_min + storage_min / area)
    l1 = (-b + np.sqrt(b * b - 4 * a * g)) / (2 * a)
    l2 = (-b - np.sqrt(b * b - 4 * a * g)) / (2 * a)
    c1, c2 = S0 - d / g, -K_p * (SP - level0) / (1 - K_d / area)
    k1, k2 = (l2 * c1 - c2) / (l2 - l1), (-l1 * c1 + c2) / (l2 - l1)
    return k1 * np.exp(l1 * t) + k2 * np.exp(l2 * t) + d / g


def test_pid_control_equation_analytic():
    # core/test/equations_test.jl:115-154: second-order linear ODE of the controlled basin
    p = build_params(MODELS["pid_control_equation"]())
    for t_end in (60.0, 300.0):
        u, S, lv, st = run_oracle(p, 1.0, t_end=t_end)
        assert S[0] == pytest.approx(_pid_equation_storage(p, t_end), rel=0.01)


def test_pid_control_tracks_the_target():
    # core/test/control_test.jl:100-131: level within 5e-3 m of the first target between days 180
    # and 330, within 5e-2 m of the moved target at the end of the run (day 700)
    p = build_params(MODELS["pid_control"]())
    target = p.nodes["pid_control"]["target"][0]
    windows = golden("pid_control_tracking")["windows"]
    for day, tol in sorted({(d, w[2]) for w in windows for d in w[:2]}):
        t = day * 86400.0
        u, S, lv, st = run_oracle(p, 3600.0, t_end=t, full_newton=1, max_halvings=20, predictor=1)
        assert st["failed"] == 0 and abs(lv[0] - target(t)) < tol


def test_discrete_control_sets_the_pid_target():
    # core/test/control_test.jl:166-199
    g = golden("discrete_control_of_pid_control")
    p = build_params(MODELS["discrete_control_of_pid_control"]())
    flat = flatten(p)
    u, S, lv, st = run_oracle(p, 3600.0, t_end=g["t_jump_days"] * 86400.0, full_newton=1, max_halvings=20,
                              predictor=1)
    assert st["failed"] == 0 and abs(lv[0] - g["target_high"]) < g["atol"]
    u, S, lv, st = run_oracle(p, 3600.0, full_newton=1, max_halvings=20, predictor=1)
    assert st["failed"] == 0 and abs(lv[0] - g["target_low"]) < g["atol"]
    rec = st["oracle"].discrete_control_record()
    assert [flat.dc_state_names[i][s_] for i, _, s_ in rec] == ["target_high", "target_low"]


@pytest.mark.parametrize("name", ["tabulated_rating_curve_control", "compound_variable_condition",
                                  "connector_node_flow_condition", "transient_condition", "storage_condition"])
def test_reference_control_sequences(name):
    # core/test/control_test.jl:133-164, 201-261, 311-332, 374-383: truth and control states of the record
    g = golden("control_records")[name]
    p = build_params(MODELS[name]())
    flat = flatten(p)
    u, S, lv, st = run_oracle(p, 3600.0, full_newton=1, max_halvings=20, predictor=1)
    assert st["failed"] == 0
    rec = st["oracle"].discrete_control_record()
    assert [flat.dc_state_names[i][s_] for i, _, s_ in rec] == g["control_state"]
    assert ["T" if b & 1 else "F" for _, b, _ in rec] == g["truth_state"]
    if "storage_below" in g:
        assert S[0] < g["storage_below"]


def test_user_demand_withdrawal_levels():
    # core/test/run_models_test.jl:388-412: the constant UserDemand withdraws down to its minimum
    # level 0.9 m (900 m3), the dynamic one to 0.5 m (500 m3)
    p = build_params(MODELS["user_demand"]())
    assert p.basin.storage0[0] == pytest.approx(1000.0)
    for days, expect, atol in golden("user_demand_levels")["checks"]:
        u, S, lv, st = run_oracle(p, 3600.0, t_end=days * 86400.0, full_newton=1, max_halvings=20, predictor=1)
        assert st["failed"] == 0 and S[0] == pytest.approx(expect, abs=atol)


def test_misc_nodes_linear_storages():
    # core/test/equations_test.jl:156-181: storage1 = S1(0) + t (q_boundary - q_pump), storage2 = S2(0) + t q_pump
    p = build_params(MODELS["misc_nodes"]())
    q_boundary, q_pump = 1.5e-4, 1e-4
    S0 = p.basin.storage0.copy()
    for days in (1, 30, 366):
        t = days * 86400.0
        u, S, lv, st = run_oracle(p, 86400.0, t_end=t)
        assert S[0] == pytest.approx(S0[0] + t * (q_boundary - q_pump), rel=1e-9)
        assert S[1] == pytest.approx(S0[1] + t * q_pump, rel=1e-9)


@pytest.mark.parametrize("name", ["basic", "backwater", "pid_control", "outlet_continuous_control",
                                  "user_demand", "misc_nodes"])
def test_dual_jacobian_matches_finite_differences(name):
    """The oracle's forward-mode Jacobian (its ForwardDiff stand-in) against central
    differences of the oracle RHS, and coloured compression == dense columns
    (core/test/differentiation_test.jl:29-55)."""
    p = build_params(MODELS[name]())
    flat = flatten(p)
    om = OracleModel(p)
    om.set_pattern(flat.jac_rows_csc, flat.jac_cols_csc)
    om.set_time_params(eval_time_params(p, 0.0))
    rng = np.random.default_rng(11)
    nb = flat["n_basin"]
    ur = np.zeros(om.n_reduced)
    ur[:nb] = p.basin.storage0 * rng.uniform(0.2, 2.0, size=nb)
    J = om.jac_dense(ur, 0.0)
    vals = om.jac_colored(ur, 0.0)
    Jc = np.zeros_like(J)
    for r, c, v in zip(flat.jac_rows_csc, flat.jac_cols_csc, vals):
        Jc[r - 1, c - 1] = v
    assert np.array_equal(J != 0, (J != 0) & (Jc != 0)), "pattern covers every non-zero"
    assert np.allclose(J, Jc, rtol=1e-14, atol=0)
    for j in range(om.n_reduced):
        h = 1e-6 * max(1.0, abs(ur[j]))
        xp, xm = ur.copy(), ur.copy()
        xp[j] += h
        xm[j] -= h
        fp, fm = om.rhs(xp, 0.0), om.rhs(xm, 0.0)
        fd = (fp - fm) / (2 * h)
        scale = np.maximum(np.abs(J[:, j]).max(), 1e-12)
        noise = 1e-15 * np.abs(fp).max() / h  # rounding of the difference quotient
        assert np.max(np.abs(fd - J[:, j])) < 1e-5 * scale + 10 * noise, (name, j)


def _standard_step_method(x, Q, w, n, h_right, h_close):
    """Backwater curve of a rectangular channel (core/test/run_models_test.jl:431-487)."""
    def friction_slope(Q, w, d, n):
        A = d * w
        R_h = A / (w + 2 * d)
        return Q ** 2 * n ** 2 / (A ** 2 * R_h ** (4 / 3))

    h = np.full(len(x), h_right)
    h1 = h_right
    for i in range(len(x) - 2, -1, -1):
        h2 = h1
        dh = h_close + 1.0
        L = x[i + 1] - x[i]
        while dh > h_close:
            h2new = h1 + 0.5 * (friction_slope(Q, w, h1, n) + friction_slope(Q, w, h2, n)) * L
            dh = abs(h2new - h2)
            h2 = h2new
        h[i] = h2
        h1 = h2
    return h


def test_backwater_against_standard_step_method():
    # core/test/run_models_test.jl:489-523: levels within 0.02 m, flow in = flow out = 5 +- 1e-3
    p = build_params(MODELS["backwater"]())
    flat = flatten(p)
    u, S, lv, st = run_oracle(p, 3600.0, full_newton=1, max_halvings=20, predictor=1)
    assert st["failed"] == 0 and st["negative"] == 0
    h_actual = lv[:50]
    x = np.arange(10.0, 991.0, 20.0)
    h_expected = _standard_step_method(x, 5.0, 1.0, 0.04, h_actual[-1], 1.0e-6)
    assert np.all(np.abs(h_expected - h_actual) < 0.02)
    om = OracleModel(p)
    om.set_time_params(eval_time_params(p, p.tables.t_end))
    du = om.rhs(om.reduce(u) * 0 + (S - p.basin.storage0), p.tables.t_end)
    assert du[flat["off_manning"] + 49] == pytest.approx(5.0, abs=1e-3)
    assert du[flat["off_manning"]] == pytest.approx(5.0, abs=1e-3)


def test_pump_discrete_control_record():
    # core/test/control_test.jl:33-44: the sequence of control state changes over the year
    p = build_params(MODELS["pump_discrete_control"]())
    flat = flatten(p)
    u, S, lv, st = run_oracle(p, 3600.0, full_newton=1, max_halvings=20, predictor=1)
    om = st["oracle"]
    rec = om.discrete_control_record()
    ids = [n.value for n in p.nodes["discrete_control"]["node_id"]]
    names = flat.dc_state_names
    g = golden("pump_discrete_control_record")
    assert [ids[i] for i, _, _ in rec] == g["control_node_id"]
    assert [names[i][s_] for i, _, s_ in rec] == g["control_state"]
    truth = ["".join("T" if (b >> j) & 1 else "F" for j in range(2 if i == 0 else 1)) for i, b, _ in rec]
    assert truth == g["truth_state"]
    # end of run: pump off and resistance inactive -> no flow (control_test.jl:62-64)
    om.set_time_params(eval_time_params(p, p.tables.t_end))
    assert st["failed"] == 0


def test_tabulated_rating_curve_model():
    # core/test/run_models_test.jl:326-352: time-varying (cyclic) rating curves
    p = build_params(MODELS["tabulated_rating_curve"]())
    trc = p.nodes["tabulated_rating_curve"]
    idx1, idx2 = trc["current_interpolation_index"]
    assert set(idx1.u) == {1}
    assert list(idx2.u) == [2, 3, 4, 2]
    assert np.allclose(idx2.t, [0.0, 2.6784e6, 5.184e6, 7.8624e6], rtol=1e-6)
    assert idx2.periodic
    assert len(trc["interpolations"]) == 5
    assert trc["interpolations"][3].t[-1] == 1.2
    u, S, lv, st = run_oracle(p, 3600.0, full_newton=1, max_halvings=20, predictor=1)
    # `≈` on Float32 goldens: rtol = sqrt(eps(Float32))
    assert np.allclose(S, [368.31558, 365.68442], rtol=3.5e-4, atol=0)
    assert st["failed"] == 0
