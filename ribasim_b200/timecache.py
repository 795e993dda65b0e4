"""Host evaluation of the time-dependent (state-independent) parameters.

In the reference these are the `TimeDependentCache` entries filled inside `water_balance!`
through `eval_time_interpolation` whenever `t` changes (core/src/util.jl:1216-1231,
core/src/parameter.jl:812-842).  They do not depend on the state, so the host evaluates them
once per RHS time and hands the small arrays to the device (or to the CPU oracle in tests).
"""

from __future__ import annotations

import math

import numpy as np

from .network import SV_CONSTANT, Params


def eval_time_params(p: Params, t: float) -> dict[str, np.ndarray]:
    n = p.nodes
    out: dict[str, np.ndarray] = {}
    out["lb_level"] = np.array([s(t) for s in n["level_boundary"]["level"]], dtype=np.float64)
    out["fb_rate_now"] = np.array([s(t) for s in n["flow_boundary"]["flow_rate"]], dtype=np.float64)

    trc = n["tabulated_rating_curve"]
    out["trc_table_idx"] = np.array([s(t) - 1 for s in trc["current_interpolation_index"]],
                                    dtype=np.int32)
    out["trc_max_downstream_level"] = np.array([s(t) for s in trc["max_downstream_level"]],
                                               dtype=np.float64)

    fd = n["flow_demand"]
    fd_index = {nid.value: i for i, nid in enumerate(fd["node_id"])}
    for key in ("pump", "outlet"):
        d = n[key]
        m = len(d["node_id"])
        fr = np.array(d["flow_rate"], dtype=np.float64).copy()
        for i, s in enumerate(d["time_dependent_flow_rate"]):
            if s is not None:
                fr[i] = s(t)
        out[f"{key}_flow_rate"] = fr
        for f in ("min_flow_rate", "max_flow_rate", "min_upstream_level", "max_downstream_level"):
            out[f"{key}_{f}"] = np.array([s(t) for s in d[f]], dtype=np.float64)
        # FlowDemand acts as a lower bound while allocation is inactive (solve.jl:719-740)
        dem = np.full(m, math.nan)
        for i, fid in enumerate(d["flow_demand_id"]):
            if fid is None:
                continue
            k = fd_index[fid.value]
            total, any_prio = 0.0, False
            for pr, has in enumerate(fd["has_demand_priority"][k]):
                if has:
                    any_prio = True
                    total += fd["demand_interpolation"][k][pr](t)
            if any_prio:
                dem[i] = total
        out[f"{key}_flow_demand"] = dem

    ud = n["user_demand"]
    n_ud, n_prio = len(ud["node_id"]), len(ud["demand_priorities"])
    demand = np.zeros((n_ud, max(n_prio, 1)))
    for i in range(n_ud):
        for k in range(n_prio):
            if not ud["has_demand_priority"][i, k]:
                continue
            # get_demand (util.jl:926-933)
            if ud["demand_from_timeseries"][i]:
                demand[i, k] = ud["demand_interpolation"][i][k](t)
            else:
                demand[i, k] = ud["demand"][i, k]
    out["ud_demand"] = demand.reshape(-1)
    out["ud_return_factor"] = np.array([s(t) for s in ud["return_factor"]], dtype=np.float64)
    # total effective demand per node (solve.jl:367-376)
    tot = np.zeros(n_ud)
    for i in range(n_ud):
        acc = 0.0
        for k in range(n_prio):
            if ud["has_demand_priority"][i, k]:
                acc += min(ud["allocated"][i, k], demand[i, k])
        tot[i] = acc
    out["ud_total_demand"] = tot

    pid = n["pid_control"]
    for f, key in (("target", "pid_target"), ("proportional", "pid_proportional"),
                   ("integral", "pid_integral"), ("derivative", "pid_derivative")):
        out[key] = np.array([s(t) for s in pid[f]], dtype=np.float64)

    consts = []
    for cvars in n["continuous_control"]["compound_variable"]:
        for sv in cvars:
            consts.append(sv["series"](t + sv["look_ahead"]) if sv["kind"] == SV_CONSTANT else 0.0)
    out["cc_sv_const"] = np.array(consts, dtype=np.float64)

    # DiscreteControl: thresholds at t and boundary sub-variables at t + look_ahead, in the
    # condition order of flatten._flatten_discrete_control
    hi, lo, dconst = [], [], []
    for cvars in n["discrete_control"]["compound_variables"]:
        for cv in cvars:
            for th, tl in zip(cv["threshold_high"], cv["threshold_low"]):
                hi.append(th(t))
                lo.append(tl(t))
                for sv in cv["subvariables"]:
                    dconst.append(sv["series"](t + sv["look_ahead"]) if sv["kind"] == SV_CONSTANT else 0.0)
    out["dc_thr_hi"] = np.array(hi, dtype=np.float64)
    out["dc_thr_lo"] = np.array(lo, dtype=np.float64)
    out["dc_sv_const"] = np.array(dconst, dtype=np.float64)
    return out


def flow_boundary_increment(p: Params, tprev: float, t: float) -> np.ndarray:
    """`integral(flow_rate, tprev, t)` per FlowBoundary (core/src/solve.jl:86-99)."""
    return np.array([s.integral(tprev, t) for s in p.nodes["flow_boundary"]["flow_rate"]],
                    dtype=np.float64)


def update_vertical_flux(p: Params, t: float) -> bool:
    """`update_basin!` (core/src/callback.jl:846-866): NaN keeps the old value.
    Returns True when something changed."""
    changed = False
    for name, series in p.basin.forcing.items():
        arr = p.basin.vertical_flux[name]
        for i, s in enumerate(series):
            v = s(t)
            if not math.isnan(v) and arr[i] != v:
                arr[i] = v
                changed = True
    return changed


def forcing_tstops(p: Params, t_end: float) -> list[float]:
    """tstops of the Basin forcing callback (core/src/callback.jl:37-43)."""
    ts: set[float] = set()
    for s in p.basin.forcing["precipitation"]:
        ts.update(s.tstops(t_end))
    return sorted(ts)
