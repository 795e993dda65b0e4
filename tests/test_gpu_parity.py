"""Parity of the CUDA path (through the C ABI of librbs.so) against the CPU oracle.

Tolerances (BASELINE.json north_star): topology/indexing bit-exact (tested on the host in
test_host_goldens.py); RHS and Jacobian entries 1e-12 relative in fp64; end-of-run storages and
cumulative flows within the solver's reltol at matched fixed-step implicit Euler.
"""

from __future__ import annotations

import numpy as np
import pytest

from tests.helpers import (FORCING, csc_vals_from_dense, csr_to_dense, gpu_handle_for, make_case,
                           oracle_for_member, rel_err, w_to_dense)

pytestmark = pytest.mark.gpu

RTOL_ENTRY = 1e-12
MODELS_RHS = ["basic", "basic_transient", "backwater", "trivial", "bucket", "linear_resistance",
              "rating_curve", "rating_curve_between_basins", "manning_resistance", "misc_nodes",
              "user_demand", "pid_control", "pid_control_equation", "outlet_continuous_control",
              "pump_discrete_control", "discrete_control_of_pid_control", "manning_discrete_control",
              "junction_combined", "junction_chained",
              "tabulated_rating_curve",
              # the bench workloads (BASELINE configs 4 and 5) at full / reduced size
              "hws_like_500", "random_tree_1500"]


def _workload(name):
    """Model tables of a named case: a reference test model, or a bench workload."""
    from ribasim_b200.testmodels import hws_like_model, random_tree_model
    if name == "hws_like_500":
        return hws_like_model(n_basins=500, n_days=2)
    if name == "random_tree_1500":
        return random_tree_model(n_basins=1500, n_hours=24)
    return name


def _scaled_err(a: np.ndarray, b: np.ndarray, axis_scale: np.ndarray) -> float:
    d = np.abs(a - b)
    d[(a == b) | (np.isnan(a) & np.isnan(b))] = 0.0
    return float(np.max(d / axis_scale)) if d.size else 0.0


@pytest.mark.parametrize("name", MODELS_RHS)
def test_rhs_jacobian_w_solve_parity(name):
    big = name in ("hws_like_500", "random_tree_1500")
    B = 3 if big else 9
    case = make_case(_workload(name), B, seed=1)
    flat, t = case["flat"], case["t"]
    h = gpu_handle_for(case)
    try:
        du_g, cache_g = h.rhs(case["ur"], t, with_cache=True)
        gamma = 3600.0
        jv, wv = h.jac(case["ur"], t, gamma=gamma)
        rng = np.random.default_rng(5)
        b = rng.normal(size=(flat["n_reduced"], B))
        x = h.solve(b)
        for m in range(B):
            om = oracle_for_member(case, m)
            du_o, cache_o = om.rhs(case["ur"][:, m], t, with_cache=True)
            # basin caches: storage is a pure sum — bit-exact against the oracle in the device's
            # association (S0 + A u) + (cumP + cumR + cumD + sum fb), and within two roundings of the
            # reference's own term-by-term order (core/src/solve.jl:187-202)
            assert np.array_equal(cache_g["storage"][:, m], cache_o["storage"])
            om_ref = oracle_for_member(case, m, merged_exact=False)
            _, cache_r = om_ref.rhs(case["ur"][:, m], t, with_cache=True)
            s_scale = np.maximum(np.abs(cache_r["storage"]), np.abs(case["p"].basin.storage0) + np.abs(case["ur"][:flat["n_basin"], m]))
            assert np.max(np.abs(cache_g["storage"][:, m] - cache_r["storage"]) / np.maximum(s_scale, 1e-300)) <= 4 * 2.3e-16, name
            for k in ("level", "area", "factor"):
                assert rel_err(cache_g[k][:, m], cache_o[k]) <= 4e-16, (name, k)
            # RHS entries: 1e-12 relative (to the entry, with a floor at 1e-12 of the vector scale)
            scale = np.maximum(np.abs(du_o), 1e-12 * max(np.abs(du_o).max(), 1e-300))
            assert _scaled_err(du_g[:, m], du_o, scale) <= RTOL_ENTRY, name
            # Jacobian entries vs the oracle's forward-mode duals: 1e-12 of the row scale
            Jo = om.jac_dense(case["ur"][:, m], t)
            Jg = csr_to_dense(flat, jv[:, m])
            assert not np.any((Jo != 0) & (Jg == 0) & (np.abs(Jo) > 0)), \
                f"{name}: analytic pattern misses a non-zero of the dual-number Jacobian"
            rs = np.maximum(np.abs(Jo).max(axis=1, keepdims=True), 1e-300)
            assert float(np.max(np.abs(Jg - Jo) / rs)) <= RTOL_ENTRY, name
            # W_inner = A J - I/gamma
            Wo = om.W_inner(csc_vals_from_dense(flat, Jo), gamma)
            Wg = w_to_dense(flat, wv[:, m])
            assert float(np.max(np.abs(Wg - Wo)) / np.abs(Wo).max()) <= RTOL_ENTRY, name
            # linear solve: backward-stable residual, forward error bounded by cond(W)
            res = np.abs(Wg @ x[:, m] - b[:, m]).max()
            assert res <= 1e-13 * (np.abs(Wg).max() * np.abs(x[:, m]).max() + np.abs(b[:, m]).max()) * \
                flat["n_reduced"], name
            if not big or m == 0:
                xo = np.linalg.solve(Wg, b[:, m])
                cond = np.linalg.cond(Wg)
                assert np.abs(x[:, m] - xo).max() <= 1e-14 * cond * np.abs(xo).max() + 1e-300, name
    finally:
        h.close()


@pytest.mark.parametrize("name", ["basic", "backwater", "user_demand", "pid_control"])
@pytest.mark.parametrize("lu_mode", [0, 1])
def test_lu_modes_agree_bitwise(name, lu_mode):
    """Level-per-launch and member-tile LU run the same schedule: identical bits."""
    B = 37
    case = make_case(name, B, seed=2)
    flat = case["flat"]
    rng = np.random.default_rng(9)
    b = rng.normal(size=(flat["n_reduced"], B))
    ref = None
    for mode in (lu_mode, 1 - lu_mode):
        h = gpu_handle_for(case, lu_mode=mode)
        try:
            h.jac(case["ur"], case["t"], gamma=600.0)
            x = h.solve(b)
        finally:
            h.close()
        if ref is None:
            ref = x
        else:
            assert np.array_equal(ref, x)


@pytest.mark.parametrize("name", ["basic", "user_demand", "misc_nodes", "outlet_continuous_control"])
def test_reduce_and_limit_flow_parity(name):
    B = 5
    case = make_case(name, B, seed=3)
    flat = case["flat"]
    n = flat["n_state"]
    rng = np.random.default_rng(4)
    uprev = np.abs(rng.normal(0, 50.0, size=(n, B)))
    u = uprev + rng.normal(0, 5.0, size=(n, B))
    h = gpu_handle_for(case)
    try:
        ur_g = h.reduce(u)
        dt = 3600.0
        lim_g = h.limit_flow(u, uprev, dt, case["t"])
        for m in range(B):
            om = oracle_for_member(case, m)
            assert np.array_equal(ur_g[:, m], om.reduce(u[:, m])), "reduce_state! is a fixed-order sum"
            # level_prev of the oracle defaults to the level at t0, like the handle's zero state
            lim_o = om.limit_flow(u[:, m], uprev[:, m], dt, case["t"])
            assert rel_err(lim_g[:, m], lim_o) <= 1e-15, name
    finally:
        h.close()


STEP_CASES = [("basic", 3600.0, 24), ("trivial", 21600.0, 12), ("user_demand", 3600.0, 12),
              ("outlet_continuous_control", 3600.0, 12), ("pid_control", 3600.0, 12),
              ("backwater", 3600.0, 6), ("hws_like", 3600.0, 6)]


@pytest.mark.parametrize("name,dt,n_steps", STEP_CASES)
@pytest.mark.parametrize("full_newton", [0, 1])
def test_implicit_euler_steps_parity(name, dt, n_steps, full_newton):
    """Device-resident sub-stepping Newton against the oracle's restatement of the same
    algorithm: same accept/reject sequence and iteration counts, states equal within Newton
    round-off (a small fraction of the solver tolerance)."""
    from oracle.oracle import _dp, advance_opts

    B = 4
    if name == "hws_like":
        from ribasim_b200.testmodels import hws_like_model
        case = make_case(hws_like_model(n_basins=60, n_days=2), B, seed=7, t=0.0)
    else:
        case = make_case(name, B, seed=7, t=0.0)
    flat = case["flat"]
    n = flat["n_state"]
    h = gpu_handle_for(case)
    try:
        h.set_solver(1e-5, 1e-5, 10)
        h.set_stepper(full_newton=bool(full_newton), predictor=True, max_halvings=20, grow_iters=4)
        opts = advance_opts(1e-5, 1e-5, 10, full_newton, 4, 20, 1, 1, 2)
        h.refresh(0.0)
        h.set_member("level_prev", h.get_member("level", flat["n_basin"]))
        oms = []
        for m in range(B):
            om = oracle_for_member(case, m)
            _, c0 = om.rhs(np.zeros(flat["n_reduced"]), 0.0, with_cache=True)
            lv = np.ascontiguousarray(c0["level"])
            om.L.orc_set_level_prev(om.h, _dp(lv))
            om.set_f64("fb_rate_step", case["fb_rate"][:, m])
            oms.append(om)
        us = [np.zeros(n) for _ in range(B)]
        etas = [1.0] * B
        htry = [dt] * B
        tot = dict(iters=0, substeps=0, rejects=0)
        t = 0.0
        for s in range(n_steps):
            st = h.step(t, dt)
            for m, om in enumerate(oms):
                rc, so, htry[m], etas[m] = om.advance(us[m], t, dt, opts, htry[m], etas[m])
                for k in tot:
                    tot[k] += so[k]
                assert (st[m] & 1) == (0 if so["converged"] else 1), (name, s, m)
            t += dt
        ug = h.get_member("u", n)
        Sg = h.get_member("storage", flat["n_basin"])
        for m, om in enumerate(oms):
            scale = 1e-5 + np.abs(us[m]) * 1e-5  # the solver's own abstol/reltol
            assert np.max(np.abs(ug[:, m] - us[m]) / scale) <= 1e-3, name
            om.set_f64("fb_incr", np.zeros(flat["n_fb"]))
            _, co = om.rhs(om.reduce(us[m]), t, with_cache=True)
            assert np.max(np.abs(Sg[:, m] - co["storage"]) /
                          (1e-5 + np.abs(co["storage"]) * 1e-5)) <= 1e-2, name
        stats = h.stats()
        assert stats["steps"] == n_steps
        assert stats["failed"] == 0
        assert (stats["substeps"], stats["newton_iters"], stats["rejects"]) == \
            (tot["substeps"], tot["iters"], tot["rejects"]), name
    finally:
        h.close()


def test_model_driver_matches_oracle_run():
    """Host driver + device stepping for 10 days of `basic` against run_oracle (same dt)."""
    from oracle.oracle import run_oracle
    from ribasim_b200.model import Model
    from ribasim_b200.network import build_params
    from ribasim_b200.testmodels import MODELS

    t_end = 10 * 86400.0
    p = build_params(MODELS["basic"]())
    u_o, S_o, lv_o, st_o = run_oracle(p, 3600.0, t_end=t_end)
    model = Model(build_params(MODELS["basic"]()), n_members=3, dt=3600.0, max_halvings=0,
                  predictor=False, full_newton=False, early_exit=False)
    try:
        model.update_until(t_end)
        S_g, u_g = model.storage(), model.u()
        for m in range(3):
            assert np.allclose(S_g[:, m], S_o, rtol=1e-5, atol=1e-5)
            assert np.allclose(u_g[:, m], u_o, rtol=1e-5, atol=1e-5)
        assert model.status_counts["negative_storage"] == 0
    finally:
        model.finalize()


def test_member_independence_and_determinism():
    """A member's result must not depend on the batch around it (that is what makes sharding
    over GPUs exact), and two runs give identical bits (no float atomics)."""
    name = "hws_like"
    from ribasim_b200.testmodels import hws_like_model
    tables = hws_like_model(n_basins=60, n_days=2)
    big = make_case(tables, 64, seed=5, t=0.0)
    flat = big["flat"]
    n = flat["n_state"]

    def run(case, members):
        sub = dict(case)
        sub["B"] = len(members)
        sub["forcing"] = {k: np.ascontiguousarray(v[:, members]) for k, v in case["forcing"].items()}
        sub["fb_rate"] = np.ascontiguousarray(case["fb_rate"][:, members])
        h = gpu_handle_for(sub)
        try:
            h.refresh(0.0)
            h.set_member("level_prev", h.get_member("level", flat["n_basin"]))
            t = 0.0
            for _ in range(6):
                h.step(t, 1800.0)
                t += 1800.0
            return h.get_member("u", n), h.get_member("storage", flat["n_basin"])
        finally:
            h.close()

    members = list(range(64))
    u_all, S_all = run(big, members)
    u_again, S_again = run(big, members)
    assert np.array_equal(u_all, u_again) and np.array_equal(S_all, S_again), name
    pick = [63, 5, 17]
    u_sub, S_sub = run(big, pick)
    assert np.array_equal(u_sub, u_all[:, pick])
    assert np.array_equal(S_sub, S_all[:, pick])


@pytest.mark.parametrize("lin_mode,n_groups", [(0, 1), (1, 1), (2, 1), (2, 3), (3, 1), (3, 3), (-1, 0)])
def test_linear_solver_tiers_agree(lin_mode, n_groups):
    """The execution tiers of the Newton linear solve (level-per-launch, factors in L2, factors
    in shared memory with the LDU schedule, record schedule with the forward substitution fused
    into the factorisation) and any member grouping give the same
    trajectories up to the rounding of their different elimination arithmetic."""
    from ribasim_b200.testmodels import hws_like_model
    case = make_case(hws_like_model(n_basins=60, n_days=2), 40, seed=13, t=0.0)
    flat = case["flat"]
    n = flat["n_state"]

    def run(**kw):
        h = gpu_handle_for(case, **kw)
        try:
            h.refresh(0.0)
            h.set_member("level_prev", h.get_member("level", flat["n_basin"]))
            st = h.advance(0.0, 3600.0, 4)
            return h.get_member("u", n), st, h.stats()
        finally:
            h.close()

    u_ref, st_ref, stats_ref = run(lin_mode=0, n_groups=1)
    u, st, stats = run(lin_mode=lin_mode, n_groups=n_groups)
    assert np.array_equal(st, st_ref) and not np.any(st & 1)
    scale = 1e-5 + np.abs(u_ref) * 1e-5
    assert np.max(np.abs(u - u_ref) / scale) <= 1e-3
    if lin_mode == 0:
        assert np.array_equal(u, u_ref)


def test_random_tree_wide_mode_matches_oracle():
    """Config-5-like network (random tree, here 1500 basins) on the level-per-launch tier that
    the 100k-basin network uses, against the oracle."""
    from oracle.oracle import _dp, advance_opts
    from ribasim_b200.testmodels import random_tree_model
    B = 3
    case = make_case(random_tree_model(n_basins=1500, n_hours=6), B, seed=17, t=0.0)
    flat = case["flat"]
    n = flat["n_state"]
    h = gpu_handle_for(case, lin_mode=0)
    try:
        h.refresh(0.0)
        h.set_member("level_prev", h.get_member("level", flat["n_basin"]))
        st = h.advance(0.0, 3600.0, 2)
        ug = h.get_member("u", n)
    finally:
        h.close()
    assert not np.any(st & 1)
    opts = advance_opts(1e-5, 1e-5, 10, 1, 4, 20, 1, 1, 2)
    for m in range(B):
        om = oracle_for_member(case, m)
        _, c0 = om.rhs(np.zeros(flat["n_reduced"]), 0.0, with_cache=True)
        lv = np.ascontiguousarray(c0["level"])
        om.L.orc_set_level_prev(om.h, _dp(lv))
        om.set_f64("fb_rate_step", case["fb_rate"][:, m])
        u = np.zeros(n)
        eta, htry = 1.0, 3600.0
        for k in range(2):
            _, _, htry, eta = om.advance(u, k * 3600.0, 3600.0, opts, htry, eta)
        scale = 1e-5 + np.abs(u) * 1e-5
        assert np.max(np.abs(ug[:, m] - u) / scale) <= 1e-2


def test_random_tree_100k_matches_oracle():
    """BASELINE config 5 AT SIZE: the 100k-basin random tree (n = 300k states, level-per-launch
    solve tier), 4 members with different forcing, 2 steps from the initial state, against the
    oracle: same accept / reject sequence per member, cumulative flows within a hundredth of the
    solver tolerance.  dt = 600 s keeps the oracle's first (transient) steps to seconds."""
    from concurrent.futures import ThreadPoolExecutor
    from oracle.oracle import _dp, advance_opts
    from ribasim_b200.testmodels import random_tree_model
    B, dt, n_steps = 4, 600.0, 2
    case = make_case(random_tree_model(n_basins=100_000, n_hours=6), B, seed=31, t=0.0)
    flat = case["flat"]
    n, nb = flat["n_state"], flat["n_basin"]
    assert nb == 100_000 and n == 300_000
    h = gpu_handle_for(case)
    try:
        h.refresh(0.0)
        h.set_member("level_prev", h.get_member("level", nb))
        st = h.advance(0.0, dt, n_steps)
        ug = h.get_member("u", n)
        stats = h.stats()
    finally:
        h.close()
    assert not np.any(st & 1)
    opts = advance_opts(1e-5, 1e-5, 10, 1, 4, 20, 1, 1, 2)

    def run_member(m):
        om = oracle_for_member(case, m)
        _, c0 = om.rhs(np.zeros(flat["n_reduced"]), 0.0, with_cache=True)
        om.L.orc_set_level_prev(om.h, _dp(np.ascontiguousarray(c0["level"])))
        om.set_f64("fb_rate_step", case["fb_rate"][:, m])
        u = np.zeros(n)
        eta, htry, tot = 1.0, dt, dict(iters=0, substeps=0, rejects=0)
        for k in range(n_steps):
            _, info, htry, eta = om.advance(u, k * dt, dt, opts, htry, eta)
            for key in tot:
                tot[key] += info[key]
        return u, tot

    with ThreadPoolExecutor(B) as ex:
        res = list(ex.map(run_member, range(B)))
    for m, (u, _) in enumerate(res):
        scale = 1e-5 + np.abs(u) * 1e-5
        assert np.max(np.abs(ug[:, m] - u) / scale) <= 1e-2, m
    # the same stepping decisions member by member
    assert stats["newton_iters"] == sum(t["iters"] for _, t in res)
    assert stats["substeps"] == sum(t["substeps"] for _, t in res)
    assert stats["rejects"] == sum(t["rejects"] for _, t in res)


def test_discrete_control_per_member():
    """DiscreteControl evaluated on the device per ensemble member: members with different forcing
    switch control states at different times; states, truth values and trajectories match the
    oracle member by member."""
    from oracle.oracle import _dp, advance_opts
    B = 6
    case = make_case("pump_discrete_control", B, seed=21, t=0.0)
    flat = case["flat"]
    n, nb = flat["n_state"], flat["n_basin"]
    # spread the members: different precipitation on the receiving basin
    case["forcing"]["precipitation"][1, :] = np.linspace(0.0, 4e-8, B)
    h = gpu_handle_for(case)
    opts = advance_opts(1e-5, 1e-5, 10, 1, 4, 20, 1, 1, 2)
    try:
        h.refresh(0.0)
        h.set_member("level_prev", h.get_member("level", nb))
        h.apply_discrete_control()
        h.refresh(0.0)
        dt, n_win, win = 86400.0, 12, 10
        oms, us, etas, htry = [], [], [1.0] * B, [dt] * B
        for m in range(B):
            om = oracle_for_member(case, m)
            _, c0 = om.rhs(np.zeros(flat["n_reduced"]), 0.0, with_cache=True)
            lv = np.ascontiguousarray(c0["level"])
            om.L.orc_set_level_prev(om.h, _dp(lv))
            om.set_f64("fb_rate_step", case["fb_rate"][:, m])
            om.apply_discrete_control(np.zeros(flat["n_reduced"]), 0.0)
            oms.append(om)
            us.append(np.zeros(n))
        st0 = h.get_member_i32("dc_state", flat["n_dc"])
        for m, om in enumerate(oms):
            assert np.array_equal(st0[:, m], om.discrete_control_state()[0])
        t = 0.0
        for k in range(n_win):
            st = h.advance(t, dt, win)
            assert not np.any(st)
            for m, om in enumerate(oms):
                for j in range(win):
                    _, _, htry[m], etas[m] = om.advance(us[m], t + j * dt, dt, opts, htry[m], etas[m])
            t += win * dt
            dcs = h.get_member_i32("dc_state", flat["n_dc"])
            dct = h.get_member_i32("dc_truth", flat["n_dc_cond"])
            for m, om in enumerate(oms):
                so, to, _, undefined = om.discrete_control_state()
                assert undefined == 0
                assert np.array_equal(dcs[:, m], so), (k, m)
                assert np.array_equal(dct[:, m], to), (k, m)
        ug = h.get_member("u", n)
        changes = h.get_member_i32("cnt_dc", 1)[0]
        for m, om in enumerate(oms):
            scale = 1e-5 + np.abs(us[m]) * 1e-5
            assert np.max(np.abs(ug[:, m] - us[m]) / scale) <= 1e-2
            assert changes[m] == om.discrete_control_state()[2]
        assert len(set(changes.tolist())) > 1 or len({tuple(c) for c in dcs.T}) > 1, \
            "members were meant to diverge"
    finally:
        h.close()


def test_manning_discrete_control_per_member():
    """ManningResistance control states (length, roughness, profile) switched per member on the
    device: members with different boundary inflow cross the threshold at different times;
    control states and trajectories match the oracle member by member."""
    from oracle.oracle import _dp, advance_opts
    B = 5
    case = make_case("manning_discrete_control", B, seed=23, t=0.0)
    flat = case["flat"]
    n, nb = flat["n_state"], flat["n_basin"]
    case["fb_rate"][0, :] = np.linspace(1.5e-3, 6e-3, B)
    h = gpu_handle_for(case)
    opts = advance_opts(1e-5, 1e-5, 10, 1, 4, 20, 1, 1, 2)
    try:
        h.refresh(0.0)
        h.set_member("level_prev", h.get_member("level", nb))
        h.apply_discrete_control()
        h.refresh(0.0)
        dt, n_win, win = 3600.0, 6, 8
        oms, us, etas, htry = [], [], [1.0] * B, [dt] * B
        for m in range(B):
            om = oracle_for_member(case, m)
            _, c0 = om.rhs(np.zeros(flat["n_reduced"]), 0.0, with_cache=True)
            om.L.orc_set_level_prev(om.h, _dp(np.ascontiguousarray(c0["level"])))
            om.set_f64("fb_rate_step", case["fb_rate"][:, m])
            om.apply_discrete_control(np.zeros(flat["n_reduced"]), 0.0)
            oms.append(om)
            us.append(np.zeros(n))
        t = 0.0
        seen = set()
        for k in range(n_win):
            st = h.advance(t, dt, win)
            assert not np.any(st)
            for m, om in enumerate(oms):
                for j in range(win):
                    _, _, htry[m], etas[m] = om.advance(us[m], t + j * dt, dt, opts, htry[m], etas[m])
            t += win * dt
            dcs = h.get_member_i32("dc_state", flat["n_dc"])
            for m, om in enumerate(oms):
                assert np.array_equal(dcs[:, m], om.discrete_control_state()[0]), (k, m)
            seen |= {tuple(c) for c in dcs.T}
        ug = h.get_member("u", n)
        for m in range(B):
            scale = 1e-5 + np.abs(us[m]) * 1e-5
            assert np.max(np.abs(ug[:, m] - us[m]) / scale) <= 1e-2
        assert len(seen) > 1, "both control states were meant to occur"
    finally:
        h.close()


def test_advance_window_equals_single_steps():
    """rbs_advance over a window = the same number of rbs_step_implicit_euler calls, bitwise,
    although members cross the outer step boundaries at different passes."""
    from ribasim_b200.testmodels import hws_like_model
    case = make_case(hws_like_model(n_basins=60, n_days=2), 48, seed=11, t=0.0)
    flat = case["flat"]
    n = flat["n_state"]
    out = []
    for windowed in (False, True):
        h = gpu_handle_for(case)
        try:
            h.set_stepper(full_newton=True)
            h.refresh(0.0)
            h.set_member("level_prev", h.get_member("level", flat["n_basin"]))
            if windowed:
                st = h.advance(0.0, 1800.0, 8)
            else:
                st = np.zeros(48, dtype=np.int32)
                for k in range(8):
                    st |= h.step(k * 1800.0, 1800.0)
            out.append((h.get_member("u", n), h.get_member("storage", flat["n_basin"]), st, h.stats()))
        finally:
            h.close()
    assert np.array_equal(out[0][0], out[1][0])
    assert np.array_equal(out[0][1], out[1][1])
    assert np.array_equal(out[0][2], out[1][2])
    for k in ("substeps", "newton_iters", "rejects"):
        assert out[0][3][k] == out[1][3][k]
    assert out[1][3]["passes"] <= out[0][3]["passes"]


def test_forced_window_equals_per_step_forcing():
    """rbs_stage_forcing_window + rbs_advance_window (forcing that changes every outer step, staged
    from pinned memory; members cross the step boundaries at their own pace) = rbs_set_member_f64 +
    rbs_step_implicit_euler per step, bitwise; the streamed-back storages are the storages at the
    end of every outer step.  Two windows back to back exercise the buffer swap."""
    import torch
    from ribasim_b200.testmodels import hws_like_model
    B, W, dt = 40, 5, 1800.0
    case = make_case(hws_like_model(n_basins=60, n_days=2), B, seed=23, t=0.0)
    flat = case["flat"]
    n, nb, nfb = flat["n_state"], flat["n_basin"], flat["n_fb"]
    rng = np.random.default_rng(5)
    keys = {"precipitation": nb, "potential_evaporation": nb, "drainage": nb, "fb_rate": nfb}
    base = {k: (case["fb_rate"] if k == "fb_rate" else case["forcing"][k]) for k in keys}
    # forcing of the 2 W outer steps: member- and step-dependent factors
    series = {k: np.stack([base[k] * rng.lognormal(0.0, 0.4, size=base[k].shape) for _ in range(2 * W)])
              for k in keys}

    def prepare():
        h = gpu_handle_for(case)
        h.set_stepper(full_newton=True)
        h.refresh(0.0)
        h.set_member("level_prev", h.get_member("level", nb))
        return h

    h = prepare()
    try:
        S_ref, U_ref, st_ref = [], [h.get_member("u", n)], np.zeros(B, dtype=np.int32)
        for s in range(2 * W):
            for k in keys:
                h.set_member(k, series[k][s])
            st_ref |= h.step(s * dt, dt)
            S_ref.append(h.get_member("storage", nb))
            U_ref.append(h.get_member("u", n))
        u_ref, stats_ref = h.get_member("u", n), h.stats()
    finally:
        h.close()
    h = prepare()
    try:
        pinned = {k: [torch.from_numpy(np.ascontiguousarray(series[k][w * W:(w + 1) * W])).pin_memory()
                      for w in range(2)] for k in keys}
        out = [torch.zeros((W, nb, B), dtype=torch.float64).pin_memory() for _ in range(2)]
        flow = [torch.zeros((W, n, B), dtype=torch.float64).pin_memory() for _ in range(2)]
        st = np.zeros(B, dtype=np.int32)
        for k in keys:
            h.stage_forcing_window(k, pinned[k][0].data_ptr(), keys[k], W)
        h.commit_forcing_window()
        for w in range(2):
            if w == 0:  # uploaded (in two chunks) while window 0 runs, live when it returns
                for k in keys:
                    sl = keys[k] * B * 8
                    h.stage_forcing_window(k, pinned[k][1].data_ptr(), keys[k], 2)
                    h.stage_forcing_window(k, pinned[k][1].data_ptr() + 2 * sl, keys[k], W - 2, first_step=2)
            st |= h.advance_window(w * W * dt, dt, W, out[w].data_ptr(), flow_out_ptr=flow[w].data_ptr())
        h.synchronize()
        u, stats = h.get_member("u", n), h.stats()
        last = {k: h.get_member(k, keys[k]) for k in keys}
    finally:
        h.close()
    assert np.array_equal(u, u_ref)
    assert np.array_equal(st, st_ref)
    for k in ("substeps", "newton_iters", "rejects"):
        assert stats[k] == stats_ref[k]
    assert stats["passes"] <= stats_ref["passes"]
    for s in range(2 * W):
        assert np.array_equal(out[s // W][s % W].numpy(), S_ref[s]), s
        # mean flow of every state over the outer step = increment of the cumulative state / dt
        assert np.array_equal(flow[s // W][s % W].numpy(), (U_ref[s + 1] - U_ref[s]) / dt), s
    for k in keys:  # the last slice stays in force
        assert np.array_equal(last[k], series[k][2 * W - 1])


def test_forcing_slicing_equals_repeated_slices():
    """rbs_set_forcing_slicing: forcing that changes every 3 outer steps, uploaded once per change,
    = the same forcing uploaded as one slice per outer step, bitwise (a window of 7 steps that
    starts at step 2 of a slice, so that the phase and a partial last slice are exercised)."""
    import torch
    from ribasim_b200.testmodels import hws_like_model
    B, W, every, phase, dt = 36, 7, 3, 2, 1800.0
    case = make_case(hws_like_model(n_basins=40, n_days=2), B, seed=29, t=0.0)
    flat = case["flat"]
    n, nb, nfb = flat["n_state"], flat["n_basin"], flat["n_fb"]
    rng = np.random.default_rng(7)
    keys = {"precipitation": nb, "potential_evaporation": nb, "drainage": nb, "fb_rate": nfb}
    base = {k: (case["fb_rate"] if k == "fb_rate" else case["forcing"][k]) for k in keys}
    n_slices = (phase + W - 1) // every + 1
    slices = {k: np.stack([base[k] * rng.lognormal(0.0, 0.4, size=base[k].shape) for _ in range(n_slices)])
              for k in keys}
    per_step = {k: np.stack([slices[k][(s + phase) // every] for s in range(W)]) for k in keys}

    def run(series, slicing):
        h = gpu_handle_for(case)
        try:
            h.set_stepper(full_newton=True)
            h.refresh(0.0)
            h.set_member("level_prev", h.get_member("level", nb))
            if slicing:
                h.set_forcing_slicing(every, phase)
            pinned = {k: torch.from_numpy(np.ascontiguousarray(series[k])).pin_memory() for k in keys}
            out = torch.zeros((W, nb, B), dtype=torch.float64).pin_memory()
            for k in keys:
                h.stage_forcing_window(k, pinned[k].data_ptr(), keys[k], series[k].shape[0])
            h.commit_forcing_window()
            st = h.advance_window(0.0, dt, W, out.data_ptr())
            h.synchronize()
            return h.get_member("u", n), st, out.numpy().copy(), {k: h.get_member(k, keys[k]) for k in keys}
        finally:
            h.close()

    u_ref, st_ref, S_ref, _ = run(per_step, False)
    u, st, S, last = run(slices, True)
    assert np.array_equal(u, u_ref) and np.array_equal(st, st_ref) and np.array_equal(S, S_ref)
    for k in keys:
        assert np.array_equal(last[k], slices[k][n_slices - 1])


def test_parameter_ensemble_per_member_overrides():
    """Parameter ensembles (north star (4)): per-member Manning n, LinearResistance resistance and
    Pump / Outlet flow rates, LevelBoundary levels.  RHS and Jacobian of every member match an oracle instance built with
    that member's parameters to 1e-12; a few implicit-Euler steps match it within the solver
    tolerance; a member whose override equals the shared value is bitwise the shared-value run."""
    from oracle.oracle import _dp, advance_opts
    from ribasim_b200.testmodels import hws_like_model
    B = 5
    case = make_case(hws_like_model(n_basins=60, n_days=2), B, seed=41, t=0.0)
    flat, tp = case["flat"], case["tp"]
    n, nb = flat["n_state"], flat["n_basin"]
    rng = np.random.default_rng(12)

    def spread(base):
        f = rng.lognormal(0.0, 0.3, size=(len(base), B))
        f[:, 0] = 1.0  # member 0 keeps the shared value
        return np.ascontiguousarray(np.asarray(base, dtype=float)[:, None] * f)

    over = {"mn_manning_n": spread(flat["mn_manning_n"]), "lr_resistance": spread(flat["lr_resistance"]),
            "pump_flow_rate": spread(tp["pump_flow_rate"]), "outlet_flow_rate": spread(tp["outlet_flow_rate"]),
            "lb_level": np.ascontiguousarray(np.asarray(tp["lb_level"])[:, None] + np.concatenate(
                [np.zeros((len(tp["lb_level"]), 1)), rng.normal(0.0, 0.2, size=(len(tp["lb_level"]), B - 1))], axis=1))}
    assert all(v.shape[0] > 0 for v in over.values())

    def oracle_member(m):
        om = oracle_for_member(case, m)
        for k, v in over.items():
            om.set_f64(k, v[:, m])
        return om

    h = gpu_handle_for(case)
    ref = gpu_handle_for(case)
    try:
        with pytest.raises(Exception, match="no per-member override"):
            h.set_member_param("mn_length", over["mn_manning_n"])
        for k, v in over.items():
            h.set_member_param(k, v)
        du, _ = h.rhs(case["ur"], 0.0, with_cache=True)
        jv, _ = h.jac(case["ur"], 0.0, gamma=3600.0)
        for m in range(B):
            om = oracle_member(m)
            du_o = om.rhs(case["ur"][:, m], 0.0)
            scale = np.maximum(np.abs(du_o), 1e-12 * np.abs(du_o).max())
            assert np.max(np.abs(du[:, m] - du_o) / scale) <= 1e-12, m
            Jo = om.jac_dense(case["ur"][:, m], 0.0)
            Jg = csr_to_dense(flat, jv[:, m])
            rs = np.maximum(np.abs(Jo).max(axis=1, keepdims=True), 1e-300)
            assert np.max(np.abs(Jg - Jo) / rs) <= 1e-12, m
        # the overrides change the answer ...
        du_ref, _ = ref.rhs(case["ur"], 0.0, with_cache=True)
        assert np.array_equal(du[:, 0], du_ref[:, 0]) and not np.array_equal(du[:, 1:], du_ref[:, 1:])
        # dropping the overrides restores the shared parameters
        for k in over:
            h.set_member_param(k, None)
        du2, _ = h.rhs(case["ur"], 0.0, with_cache=True)
        assert np.array_equal(du2, du_ref)
        for k, v in over.items():
            h.set_member_param(k, v)
        # ... and a few steps follow the per-member oracle
        for x in (h, ref):
            x.refresh(0.0)
            x.set_member("level_prev", x.get_member("level", nb))
        dt = 3600.0
        st = h.advance(0.0, dt, 3)
        ref.advance(0.0, dt, 3)
        assert not np.any(st & 1)
        ug = h.get_member("u", n)
        assert np.array_equal(ug[:, 0], ref.get_member("u", n)[:, 0])
        opts = advance_opts(1e-5, 1e-5, 10, 1, 4, 20, 1, 1, 2)
        for m in (1, B - 1):
            om = oracle_member(m)
            _, c0 = om.rhs(np.zeros(flat["n_reduced"]), 0.0, with_cache=True)
            om.L.orc_set_level_prev(om.h, _dp(np.ascontiguousarray(c0["level"])))
            om.set_f64("fb_rate_step", case["fb_rate"][:, m])
            u = np.zeros(n)
            eta, htry = 1.0, dt
            for k in range(3):
                _, _, htry, eta = om.advance(u, k * dt, dt, opts, htry, eta)
            scale = 1e-5 + np.abs(u) * 1e-5
            assert np.max(np.abs(ug[:, m] - u) / scale) <= 1e-2, m
    finally:
        h.close()
        ref.close()


def test_async_member_staging_matches_synchronous_calls():
    """rbs_set_member_f64_async / rbs_get_member_f64_async (pinned buffers, copy stream): values
    staged before step k apply from step k + 1 on, the background download returns the array as
    it was when requested; trajectories are bitwise those of the synchronous calls."""
    from ribasim_b200 import lib
    B = 12
    case = make_case("basic", B, seed=19, t=0.0)
    flat = case["flat"]
    nb, n = flat["n_basin"], flat["n_state"]
    rng = np.random.default_rng(2)
    series = [case["forcing"]["precipitation"] * rng.lognormal(0.0, 0.5, size=(nb, B)) for _ in range(5)]
    dt = 3600.0

    def prepare():
        h = gpu_handle_for(case)
        h.refresh(0.0)
        h.set_member("level_prev", h.get_member("level", nb))
        h.set_member("precipitation", series[0])
        return h

    ref = prepare()
    try:
        S_ref = []
        for k in range(4):
            ref.step(k * dt, dt)
            S_ref.append(ref.get_member("storage", nb))
            ref.set_member("precipitation", series[k + 1])
        u_ref = ref.get_member("u", n)
    finally:
        ref.close()
    h = prepare()
    try:
        stage = [lib.PinnedArray((nb, B)) for _ in range(2)]
        out = [lib.PinnedArray((nb, B)) for _ in range(4)]
        for k in range(4):
            stage[k % 2].array[:] = series[k + 1]
            h.set_member_ptr_async("precipitation", stage[k % 2].ptr, nb)  # live after step k
            h.step(k * dt, dt)
            h.get_member_ptr_async("storage", out[k].ptr, nb)
        h.synchronize()
        for k in range(4):
            assert np.array_equal(out[k].array, S_ref[k]), k
        assert np.array_equal(h.get_member("u", n), u_ref)
        for a in stage + out:
            a.close()
    finally:
        h.close()


@pytest.mark.parametrize("name", ["basic", "backwater", "user_demand", "pid_control",
                                  "outlet_continuous_control", "pump_discrete_control", "hws_like"])
def test_cooperative_window_kernel_equals_stream_path(name):
    """Small ensembles run a window as one cooperative kernel (rbs_coop.cuh: grid barriers between
    the phases, predicates instead of work lists).  Same device functions, same summation orders:
    states, storages, status flags and the iteration / sub-step / reject counters are bitwise those
    of the stream path (one launch per phase, compacted member lists)."""
    from ribasim_b200.testmodels import hws_like_model
    B = 37
    case = make_case(hws_like_model(n_basins=60, n_days=2) if name == "hws_like" else name, B, seed=29, t=0.0)
    flat = case["flat"]
    n, nb = flat["n_state"], flat["n_basin"]
    out = []
    for coop in (0, 4096):
        h = gpu_handle_for(case, coop_members=coop)
        try:
            h.refresh(0.0)
            h.set_member("level_prev", h.get_member("level", nb))
            if flat["n_dc"] > 0:
                h.apply_discrete_control()
                h.refresh(0.0)
            l0 = h.kernel_launches()
            st = h.advance(0.0, 3600.0, 5)
            st |= h.step(5 * 3600.0, 1800.0)
            launches = h.kernel_launches() - l0
            out.append((h.get_member("u", n), h.get_member("storage", nb), st, h.stats(), launches,
                        h.get_member_i32("dc_state", flat["n_dc"]) if flat["n_dc"] else None))
        finally:
            h.close()
    (u0, S0, st0, stats0, l_stream, dc0), (u1, S1, st1, stats1, l_coop, dc1) = out
    assert np.array_equal(u0, u1) and np.array_equal(S0, S1) and np.array_equal(st0, st1)
    for k in ("substeps", "newton_iters", "rejects", "failed"):
        assert stats0[k] == stats1[k], k
    if dc0 is not None:
        assert np.array_equal(dc0, dc1)
    assert l_coop <= 4 < l_stream  # k_step_begin + the cooperative kernel per call


@pytest.mark.parametrize("tile", [2, 4])
@pytest.mark.parametrize("name", ["basic", "backwater", "user_demand", "pid_control",
                                  "outlet_continuous_control", "pump_discrete_control", "hws_like"])
def test_tile_window_kernel_equals_stream_path(name, tile):
    """A block per member tile runs the whole window (rbs_tile.cuh: block barriers between the
    phases, intermediates stay in L1 / L2).  Same device functions, same summation orders: states,
    storages, status flags and the iteration / sub-step / reject counters are bitwise those of the
    stream path, also for a member count that does not fill the last tile."""
    from ribasim_b200.testmodels import hws_like_model
    B = 37
    case = make_case(hws_like_model(n_basins=60, n_days=2) if name == "hws_like" else name, B, seed=31, t=0.0)
    flat = case["flat"]
    n, nb = flat["n_state"], flat["n_basin"]
    out = []
    for tw in (0, tile):
        h = gpu_handle_for(case, coop_members=0, tile_window=tw)
        try:
            h.refresh(0.0)
            h.set_member("level_prev", h.get_member("level", nb))
            if flat["n_dc"] > 0:
                h.apply_discrete_control()
                h.refresh(0.0)
            l0 = h.kernel_launches()
            st = h.advance(0.0, 3600.0, 5)
            st |= h.step(5 * 3600.0, 1800.0)
            launches = h.kernel_launches() - l0
            out.append((h.get_member("u", n), h.get_member("storage", nb), st, h.stats(), launches,
                        h.get_member_i32("dc_state", flat["n_dc"]) if flat["n_dc"] else None))
        finally:
            h.close()
    (u0, S0, st0, stats0, l_stream, dc0), (u1, S1, st1, stats1, l_tile, dc1) = out
    assert np.array_equal(u0, u1) and np.array_equal(S0, S1) and np.array_equal(st0, st1)
    for k in ("substeps", "newton_iters", "rejects", "failed"):
        assert stats0[k] == stats1[k], k
    if dc0 is not None:
        assert np.array_equal(dc0, dc1)
    assert l_tile <= 4 < l_stream  # k_step_begin + the tile kernel per call


@pytest.mark.parametrize("name", ["basic", "backwater", "pid_control", "pump_discrete_control", "hws_like"])
def test_compact_solve_slots_equal_plain(name, monkeypatch):
    """The full-Newton path solves on compacted shared-memory slots (fill-ins take over the slots
    of dead multipliers, symbolic._compact_slots): the same arithmetic in other places, so the
    trajectories are bitwise those of the plain slot layout (`RBS_ENT_COMPACT=0`)."""
    from ribasim_b200.testmodels import hws_like_model
    B = 21
    case = make_case(hws_like_model(n_basins=120, n_days=2) if name == "hws_like" else name, B, seed=37, t=0.0)
    flat = case["flat"]
    n, nb = flat["n_state"], flat["n_basin"]
    out = []
    for compact in ("0", "1"):
        monkeypatch.setenv("RBS_ENT_COMPACT", compact)
        h = gpu_handle_for(case, coop_members=0)
        try:
            h.refresh(0.0)
            h.set_member("level_prev", h.get_member("level", nb))
            if flat["n_dc"] > 0:
                h.apply_discrete_control()
                h.refresh(0.0)
            st = h.advance(0.0, 3600.0, 6)
            out.append((h.get_member("u", n), h.get_member("storage", nb), st, h.stats()))
        finally:
            h.close()
    (u0, S0, st0, s0), (u1, S1, st1, s1) = out
    assert np.array_equal(u0, u1) and np.array_equal(S0, S1) and np.array_equal(st0, st1)
    for k in ("substeps", "newton_iters", "rejects", "failed"):
        assert s0[k] == s1[k], k


@pytest.mark.parametrize("full_newton", [1, 0])
@pytest.mark.parametrize("warps", [32, 16, 8])
@pytest.mark.parametrize("name", ["basic", "backwater", "pid_control", "pump_discrete_control", "hws_like"])
def test_lane_solve_equals_shared_memory_solve(name, warps, full_newton, monkeypatch):
    """Member-per-lane solve on global slots (rbs_lane.cuh, `RBS_LIN_LANE` = warps per block): the
    entry schedule of the shared-memory kernel with the mapping turned around (lane = member, warp =
    entry, team partial sums as per-lane accumulators combined in the butterfly's order), so the
    trajectories are bitwise those of k_newton_linear_ent — with full Newton and with frozen factors
    (members that keep their factors only update the right-hand-side column); 70 members = two
    full blocks of 32 and a ragged one."""
    from ribasim_b200.testmodels import hws_like_model
    B = 70
    case = make_case(hws_like_model(n_basins=120, n_days=2) if name == "hws_like" else name, B, seed=41, t=0.0)
    flat = case["flat"]
    n, nb = flat["n_state"], flat["n_basin"]
    out = []
    monkeypatch.setenv("RBS_LIN_LANE_MIN", "0")  # by default only launches of >= 2048 members take it
    for lane in ("0", str(warps)):
        monkeypatch.setenv("RBS_LIN_LANE", lane)
        h = gpu_handle_for(case, coop_members=0)
        try:
            h.set_stepper(full_newton=bool(full_newton))
            h.refresh(0.0)
            h.set_member("level_prev", h.get_member("level", nb))
            if flat["n_dc"] > 0:
                h.apply_discrete_control()
                h.refresh(0.0)
            st = h.advance(0.0, 3600.0, 6)
            out.append((h.get_member("u", n), h.get_member("storage", nb), st, h.stats()))
        finally:
            h.close()
    (u0, S0, st0, s0), (u1, S1, st1, s1) = out
    assert np.array_equal(u0, u1) and np.array_equal(S0, S1) and np.array_equal(st0, st1)
    for k in ("substeps", "newton_iters", "rejects", "failed"):
        assert s0[k] == s1[k], k


def test_lane_solve_taken_by_large_launches(monkeypatch):
    """Default selection: a launch of at least 2048 members (one member group of 2100 here) takes the
    member-per-lane solve, smaller ones the shared-memory kernel; members finish at different passes,
    so both kernels serve the same window — and the trajectories are bitwise those of a run with
    the lane kernel switched off."""
    B = 2100
    case = make_case("basic", B, seed=43, t=0.0)
    flat = case["flat"]
    n, nb = flat["n_state"], flat["n_basin"]
    out = []
    for lane in ("0", None):
        if lane is None:
            monkeypatch.delenv("RBS_LIN_LANE", raising=False)
        else:
            monkeypatch.setenv("RBS_LIN_LANE", lane)
        h = gpu_handle_for(case, coop_members=0, n_groups=1)
        try:
            h.refresh(0.0)
            h.set_member("level_prev", h.get_member("level", nb))
            h.profile(True)
            st = h.advance(0.0, 3600.0, 4)
            rep = h.profile_report()
            out.append((h.get_member("u", n), h.get_member("storage", nb), st, h.stats(), rep))
        finally:
            h.close()
    (u0, S0, st0, s0, r0), (u1, S1, st1, s1, r1) = out
    assert "k_newton_linear_lane" not in r0 and r1.get("k_newton_linear_lane", (0, 0.0))[0] > 0
    assert np.array_equal(u0, u1) and np.array_equal(S0, S1) and np.array_equal(st0, st1)
    for k in ("substeps", "newton_iters", "rejects", "failed"):
        assert s0[k] == s1[k], k


@pytest.mark.parametrize("cluster", [1, 4, 8])
@pytest.mark.parametrize("name", ["basic", "backwater", "user_demand", "pid_control",
                                  "outlet_continuous_control", "pump_discrete_control", "hws_like"])
def test_cluster_finalize_pass_equals_round1_launch_sequence(name, cluster):
    """Round-2 pass: the convergence test and the whole finalize half run as ONE thread-block-cluster
    kernel at the end of the pass (rbs_fused.cuh k_finalize, cluster barriers between the phases)
    instead of six launches on a compacted list at the start of the next pass.  Same device
    functions in the same per-member order: states, storages, status flags, control states and the
    iteration / sub-step / reject counters are bitwise those of the round-1 sequence; 70 members =
    two full 32-member tiles and a ragged one."""
    from ribasim_b200.testmodels import hws_like_model
    B = 70
    case = make_case(hws_like_model(n_basins=60, n_days=2) if name == "hws_like" else name, B, seed=43, t=0.0)
    flat = case["flat"]
    n, nb = flat["n_state"], flat["n_basin"]
    out = []
    for r2 in (0, 1):
        h = gpu_handle_for(case, coop_members=0, lin_m=0, r2_pass=r2, fin_cluster=cluster)
        try:
            h.refresh(0.0)
            h.set_member("level_prev", h.get_member("level", nb))
            if flat["n_dc"] > 0:
                h.apply_discrete_control()
                h.refresh(0.0)
            l0 = h.kernel_launches()
            st = h.advance(0.0, 3600.0, 5)
            st |= h.step(5 * 3600.0, 1800.0)
            launches = h.kernel_launches() - l0
            out.append((h.get_member("u", n), h.get_member("storage", nb), st, h.stats(), launches,
                        h.get_member_i32("dc_state", flat["n_dc"]) if flat["n_dc"] else None))
        finally:
            h.close()
    (u0, S0, st0, stats0, l_r1, dc0), (u1, S1, st1, stats1, l_r2, dc1) = out
    assert np.array_equal(u0, u1) and np.array_equal(S0, S1) and np.array_equal(st0, st1)
    for k in ("substeps", "newton_iters", "rejects", "failed"):
        assert stats0[k] == stats1[k], k
    if dc0 is not None:
        assert np.array_equal(dc0, dc1)
    assert l_r2 < l_r1


@pytest.mark.parametrize("B", [37, 40])
@pytest.mark.parametrize("lin_m", [4, 8])
@pytest.mark.parametrize("name", ["basic", "backwater", "pid_control", "pump_discrete_control", "hws_like"])
def test_m_member_solve_equals_two_member_tiles(name, lin_m, B):
    """The full-Newton solve with M = 4 / 8 members per thread (rbs_fused.cuh) runs the schedule of
    the 2-member tiles with the same products, summation order and team butterflies: trajectories
    are bitwise equal, with aligned member pairs (B = 40: 16-byte copies) and without (B = 37)."""
    from ribasim_b200.testmodels import hws_like_model
    case = make_case(hws_like_model(n_basins=120, n_days=2) if name == "hws_like" else name, B, seed=41, t=0.0)
    flat = case["flat"]
    n, nb = flat["n_state"], flat["n_basin"]
    out = []
    for m in (0, lin_m):
        h = gpu_handle_for(case, coop_members=0, lin_m=m)
        try:
            h.refresh(0.0)
            h.set_member("level_prev", h.get_member("level", nb))
            if flat["n_dc"] > 0:
                h.apply_discrete_control()
                h.refresh(0.0)
            st = h.advance(0.0, 3600.0, 6)
            out.append((h.get_member("u", n), h.get_member("storage", nb), st, h.stats()))
        finally:
            h.close()
    (u0, S0, st0, s0), (u1, S1, st1, s1) = out
    assert np.array_equal(u0, u1) and np.array_equal(S0, S1) and np.array_equal(st0, st1)
    for k in ("substeps", "newton_iters", "rejects", "failed"):
        assert s0[k] == s1[k], k


def test_forced_window_argument_errors():
    """The forced-window entry points fail loudly: unknown forcing array, size mismatch, fewer
    slices staged than steps advanced; a window without staged slices is a plain rbs_advance."""
    from ribasim_b200 import lib
    case = make_case("basic", 4, seed=3, t=0.0)
    flat = case["flat"]
    nb = flat["n_basin"]
    ref = gpu_handle_for(case)
    h = gpu_handle_for(case)
    try:
        for x in (ref, h):
            x.refresh(0.0)
            x.set_member("level_prev", x.get_member("level", nb))
        buf = lib.PinnedArray((3, nb, 4))  # page-locked memory from the library (rbs_host_alloc)
        buf.array[:] = case["forcing"]["precipitation"]
        with pytest.raises(lib.RbsError, match="not a forcing array"):
            h.stage_forcing_window("level", buf.ptr, nb, 3)
        with pytest.raises(lib.RbsError, match="size mismatch"):
            h.stage_forcing_window("precipitation", buf.ptr, nb + 1, 3)
        h.stage_forcing_window("precipitation", buf.ptr, nb, 3)
        h.commit_forcing_window()
        with pytest.raises(lib.RbsError, match="fewer steps staged"):
            h.advance_window(0.0, 3600.0, 4)
        st = h.advance_window(0.0, 3600.0, 3)      # the staged slices equal the current forcing
        st2 = h.advance_window(3 * 3600.0, 3600.0, 2)  # nothing staged: constant forcing
        st_ref = ref.advance(0.0, 3600.0, 5)
        assert np.array_equal(st | st2, st_ref)
        assert np.array_equal(h.get_member("u", flat["n_state"]), ref.get_member("u", flat["n_state"]))
        h.synchronize()
        buf.close()
    finally:
        h.close()
        ref.close()


@pytest.mark.parametrize("n_basins,B", [(200, 512), (500, 4096)])
def test_water_balance_closes_at_full_size(n_basins, B):
    """Size-independent property up to the full bench size (BASELINE config 4: 500 basins x 4096
    members): after k steps every basin satisfies S = S0 + A u + exact inflows (the identity
    reduce_state! + formulate_storages! encode), nothing failed, no storage went negative."""
    from ribasim_b200.testmodels import hws_like_model
    tables = hws_like_model(n_basins=n_basins, n_days=2)
    case = make_case(tables, B, seed=8, t=0.0)
    flat = case["flat"]
    nb, n = flat["n_basin"], flat["n_state"]
    h = gpu_handle_for(case)
    try:
        h.refresh(0.0)
        h.set_member("level_prev", h.get_member("level", nb))
        t, dt = 0.0, 3600.0
        st = h.advance(t, dt, 4)
        t += 4 * dt
        assert not np.any(st), "forced accepts / negative storage"
        u = h.get_member("u", n)
        S = h.get_member("storage", nb)
        assert S.min() >= 0.0
        ur = h.reduce(u)[:nb]
        p = case["p"]
        exact = case["forcing"]["precipitation"] * np.array([a[-1] for a in p.basin.areas])[:, None] * t \
            + case["forcing"]["surface_runoff"] * t + case["forcing"]["drainage"] * t
        for k in range(flat["n_fb"]):
            for b in range(nb):
                lo, hi = flat["bfb_ptr"][b], flat["bfb_ptr"][b + 1]
                if k in flat["bfb_idx"][lo:hi]:
                    exact[b] += case["fb_rate"][k] * t
        S_expect = p.basin.storage0[:, None] + ur + exact
        scale = np.maximum(np.abs(S_expect), 1.0)
        assert np.max(np.abs(S - S_expect) / scale) < 1e-11
    finally:
        h.close()


@pytest.mark.parametrize("n_basins,B", [(200, 512), (500, 4096)])
def test_forced_window_outputs_close_the_water_balance(n_basins, B):
    """Size-independent property of the streamed window outputs up to the full bench size: for
    every basin, member and outer step the storage increment equals dt x (the mean flows of the
    incident states with their signs - evaporation - infiltration + the step's own forcing slice),
    i.e. storages, mean flows and forcing slices belong to the same outer step."""
    import torch
    from ribasim_b200.testmodels import hws_like_model
    W, dt = 4, 3600.0
    case = make_case(hws_like_model(n_basins=n_basins, n_days=2), B, seed=31, t=0.0)
    flat, p = case["flat"], case["p"]
    nb, n, nfb = flat["n_basin"], flat["n_state"], flat["n_fb"]
    rng = np.random.default_rng(9)
    keys = {"precipitation": nb, "drainage": nb, "fb_rate": nfb}
    base = {k: (case["fb_rate"] if k == "fb_rate" else case["forcing"][k]) for k in keys}
    fac = rng.lognormal(0.0, 0.3, size=(W, 1, 1))  # one factor per outer step keeps memory small
    pinned = {k: torch.from_numpy(np.ascontiguousarray(base[k][None] * fac)).pin_memory() for k in keys}
    out = torch.zeros((W, nb, B), dtype=torch.float64).pin_memory()
    flow = torch.zeros((W, n, B), dtype=torch.float64).pin_memory()
    h = gpu_handle_for(case)
    try:
        h.refresh(0.0)
        h.set_member("level_prev", h.get_member("level", nb))
        S0 = h.get_member("storage", nb)
        for k in keys:
            h.stage_forcing_window(k, pinned[k].data_ptr(), keys[k], W)
        h.commit_forcing_window()
        st = h.advance_window(0.0, dt, W, out.data_ptr(), flow_out_ptr=flow.data_ptr())
        h.synchronize()
        assert not np.any(st), "forced accepts / negative storage"
    finally:
        h.close()
    S = np.concatenate([S0[None], out.numpy()])
    F = flow.numpy()
    area = np.array([a[-1] for a in p.basin.areas])[:, None]
    inc_ptr, inc_state, inc_sign = flat["inc_ptr"], flat["inc_state"], flat["inc_sign"]
    worst = 0.0
    for k in range(W):
        net = np.zeros((nb, B))
        for b in range(nb):
            for q in range(inc_ptr[b], inc_ptr[b + 1]):
                net[b] += inc_sign[q] * F[k, inc_state[q]]
            for q in range(flat["bfb_ptr"][b], flat["bfb_ptr"][b + 1]):
                net[b] += pinned["fb_rate"][k, flat["bfb_idx"][q]].numpy()
        net -= F[k, flat["off_evap"]:flat["off_evap"] + nb] + F[k, flat["off_infil"]:flat["off_infil"] + nb]
        net += pinned["precipitation"][k].numpy() * area + case["forcing"]["surface_runoff"] \
            + pinned["drainage"][k].numpy()
        dS = S[k + 1] - S[k]
        scale = np.maximum(np.abs(S[k + 1]), 1.0)
        worst = max(worst, float(np.max(np.abs(dS - dt * net) / scale)))
    assert worst < 1e-10, worst


def test_full_size_ensemble_members_match_oracle():
    """BASELINE config 4 at full size (500 basins x 4096 members, four member groups): sampled
    members of the ensemble follow the oracle's restatement of the same stepping algorithm within
    a small fraction of the solver tolerance."""
    from oracle.oracle import _dp, advance_opts
    from ribasim_b200.testmodels import hws_like_model
    B = 4096
    case = make_case(hws_like_model(n_basins=500, n_days=2), B, seed=77, t=0.0)
    flat = case["flat"]
    n, nb = flat["n_state"], flat["n_basin"]
    h = gpu_handle_for(case)
    try:
        h.refresh(0.0)
        h.set_member("level_prev", h.get_member("level", nb))
        n_steps = 24
        st = h.advance(0.0, 3600.0, n_steps)
        ug = h.get_member("u", n)
        Sg = h.get_member("storage", nb)
    finally:
        h.close()
    assert not np.any(st & 1)
    # 32 members: first / last of every member group, and interior ones
    members = sorted({0, 1023, 1024, 2047, 2048, 3071, 3072, B - 1} | set(range(7, B, 171)))[:32]
    assert len(members) == 32
    opts = advance_opts(1e-5, 1e-5, 10, 1, 4, 20, 1, 1, 2)

    def run_member(m):
        om = oracle_for_member(case, m)
        _, c0 = om.rhs(np.zeros(flat["n_reduced"]), 0.0, with_cache=True)
        om.L.orc_set_level_prev(om.h, _dp(np.ascontiguousarray(c0["level"])))
        om.set_f64("fb_rate_step", case["fb_rate"][:, m])
        u = np.zeros(n)
        eta, htry = 1.0, 3600.0
        for k in range(n_steps):
            _, _, htry, eta = om.advance(u, k * 3600.0, 3600.0, opts, htry, eta)
        _, c1 = om.rhs(om.reduce(u), n_steps * 3600.0, with_cache=True)
        return m, u, c1["storage"]

    from concurrent.futures import ThreadPoolExecutor
    with ThreadPoolExecutor(8) as ex:  # the oracle's ctypes calls release the GIL
        for m, u, S in ex.map(run_member, members):
            # end-of-run cumulative flows and storages within the solver's reltol (north star),
            # here to a hundredth of it
            scale = 1e-5 + np.abs(u) * 1e-5
            assert np.max(np.abs(ug[:, m] - u) / scale) <= 1e-2, m
            assert np.max(np.abs(Sg[:, m] - S) / (1e-5 + np.abs(S) * 1e-5)) <= 1e-2, m


def _nccl():
    import ctypes
    for name in ("libnccl.so.2", "libnccl.so"):
        try:
            return ctypes.CDLL(name, mode=ctypes.RTLD_GLOBAL)
        except OSError:
            continue
    pytest.skip("NCCL library not found")


def test_rbs_gather_all_gathers_member_arrays():
    """rbs_gather (the design's only collective) over the ranks of an NCCL communicator
    created by the host: one rank per visible GPU of this process (a single rank on a one-GPU box:
    the all-gather is then the identity).  Every rank receives the rank blocks in rank order."""
    import ctypes
    import threading

    import torch
    nccl = _nccl()
    n_dev = min(torch.cuda.device_count(), 2)
    comms = (ctypes.c_void_p * n_dev)()
    devs = (ctypes.c_int * n_dev)(*range(n_dev))
    assert nccl.ncclCommInitAll(comms, n_dev, devs) == 0
    B = 5
    case = make_case("basic", B, seed=11, t=0.0)
    flat = case["flat"]
    nb = flat["n_basin"]
    from ribasim_b200 import lib
    handles, recv, local = [], [], []
    try:
        for d in range(n_dev):
            h = lib.Handle(flat, n_members=B, device=d)
            h.set_time_params(case["tp"])
            h.set_member("precipitation", case["forcing"]["precipitation"] * (d + 1))
            h.refresh(0.0)
            h.set_member("level_prev", h.get_member("level", nb))
            h.advance(0.0, 3600.0, 2)
            handles.append(h)
            local.append(h.get_member("storage", nb))
            recv.append(torch.zeros((n_dev, nb, B), dtype=torch.float64, device=f"cuda:{d}"))
        errs = []

        def run(d):
            try:
                handles[d].gather(comms[d], "storage", nb, recv[d].data_ptr())
            except Exception as e:  # noqa: BLE001
                errs.append(e)

        threads = [threading.Thread(target=run, args=(d,)) for d in range(n_dev)]
        for t in threads:
            t.start()
        for t in threads:
            t.join()
        assert not errs, errs
        for d in range(n_dev):
            got = recv[d].cpu().numpy()
            for r in range(n_dev):
                assert np.array_equal(got[r], local[r]), (d, r)
        with pytest.raises(lib.RbsError, match="size mismatch"):
            handles[0].gather(comms[0], "storage", nb + 1, recv[0].data_ptr())
    finally:
        for h in handles:
            h.close()
        for d in range(n_dev):
            nccl.ncclCommDestroy(ctypes.c_void_p(comms[d]))
