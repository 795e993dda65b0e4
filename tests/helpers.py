"""Shared helpers of the parity tests: build the same member states / forcing for the CUDA
library and for the CPU oracle."""

from __future__ import annotations

import numpy as np

from oracle.oracle import OracleModel
from ribasim_b200.flatten import flatten
from ribasim_b200.network import build_params
from ribasim_b200.testmodels import MODELS
from ribasim_b200.timecache import eval_time_params, flow_boundary_increment

def golden(key: str) -> dict:
    """Known answers of the reference's tests (tests/golden/reference_goldens.json)."""
    import json
    from pathlib import Path

    return json.loads((Path(__file__).parent / "golden" / "reference_goldens.json").read_text())[key]


FORCING = ("precipitation", "surface_runoff", "potential_evaporation", "drainage", "infiltration")


def make_case(name: str, B: int, seed: int = 0, t: float = 3600.0, **kw):
    """Params, flat arrays and per-member random reduced states / forcing for model `name`."""
    tables = MODELS[name](**kw) if isinstance(name, str) else name
    p = build_params(tables)
    flat = flatten(p)
    rng = np.random.default_rng(seed)
    nb, n_r, n = flat["n_basin"], flat["n_reduced"], flat["n_state"]
    s0 = p.basin.storage0
    thr = p.basin.low_storage_threshold
    ur = np.zeros((n_r, B))
    # span dry, low-storage, normal and (slightly) negative storages
    scale = rng.choice([-1.02, -0.999, -0.9, -0.5, 0.0, 0.5, 3.0], size=(nb, B))
    ur[:nb] = s0[:, None] * scale + rng.uniform(0, 1.5, size=(nb, B)) * thr[:, None]
    ur[:nb, 0] = 0.0  # member 0 = the initial state
    ur[nb:] = rng.normal(0, 10.0, size=(n_r - nb, B))
    forcing = {}
    for k in FORCING:
        base = p.basin.vertical_flux[k]
        f = base[:, None] * rng.lognormal(0.0, 0.3, size=(nb, B))
        if k == "infiltration":
            f = f + rng.uniform(0, 1e-6, size=(nb, B))
        forcing[k] = np.ascontiguousarray(f)
    nfb = flat["n_fb"]
    fb_base = np.array([s(0.0) for s in p.nodes["flow_boundary"]["flow_rate"]])
    fb_rate = np.ascontiguousarray(fb_base[:, None] * rng.lognormal(0.0, 0.2, size=(nfb, B))) \
        if nfb else np.zeros((0, B))
    tp = eval_time_params(p, t)
    return dict(p=p, flat=flat, ur=ur, forcing=forcing, fb_rate=fb_rate, tp=tp, t=t, B=B)


def oracle_for_member(case, m: int, merged_exact: bool = True, tprev: float = 0.0) -> OracleModel:
    p, flat = case["p"], case["flat"]
    om = OracleModel(p)
    om.set_pattern(flat.jac_rows_csc, flat.jac_cols_csc)
    om.L.orc_set_merged_exact(om.h, int(merged_exact))
    om.set_time_params(case["tp"])
    for k in FORCING:
        om.set_f64(k, case["forcing"][k][:, m])
    om.set_f64("fb_incr", case["fb_rate"][:, m] * (case["t"] - tprev))
    om.set_f64("fb_rate_now", case["fb_rate"][:, m])  # instantaneous rate seen by the PID D-term
    om.L.orc_set_tprev(om.h, tprev)
    return om


def gpu_handle_for(case, **kw):
    from ribasim_b200 import lib

    h = lib.Handle(case["flat"], n_members=case["B"], **kw)
    h.set_time_params(case["tp"])
    for k in FORCING:
        h.set_member(k, case["forcing"][k])
    if case["flat"]["n_fb"]:
        h.set_member("fb_rate", case["fb_rate"])
    ud = case["p"].nodes["user_demand"]
    h.set_param("ud_demand_from_timeseries", np.asarray(ud["demand_from_timeseries"], dtype=np.int32))
    if len(ud["node_id"]):
        h.set_param("ud_alloc_total", np.minimum(ud["demand"], ud["allocated"]).sum(axis=1))
    h.set_tprev(0.0)
    return h


def csr_to_dense(flat, vals: np.ndarray) -> np.ndarray:
    """J_intermediate values in CSR order -> dense (n_state, n_reduced)."""
    n, n_r = flat["n_state"], flat["n_reduced"]
    J = np.zeros((n, n_r))
    rp, col = flat["jrow_ptr"], flat["jcol"]
    for r in range(n):
        for e in range(rp[r], rp[r + 1]):
            J[r, col[e]] = vals[e]
    return J


def w_to_dense(flat, vals: np.ndarray) -> np.ndarray:
    n_r = flat["n_reduced"]
    W = np.zeros((n_r, n_r))
    rp, col = flat["wrow_ptr"], flat["wcol"]
    for i in range(n_r):
        for e in range(rp[i], rp[i + 1]):
            W[i, col[e]] = vals[e]
    return W


def csc_vals_from_dense(flat, J: np.ndarray) -> np.ndarray:
    return np.array([J[r - 1, c - 1] for r, c in zip(flat.jac_rows_csc, flat.jac_cols_csc)])


def rel_err(a: np.ndarray, b: np.ndarray) -> float:
    a, b = np.asarray(a, float), np.asarray(b, float)
    both_nan = np.isnan(a) & np.isnan(b)
    both_inf = np.isinf(a) & np.isinf(b) & (np.sign(a) == np.sign(b))
    d = np.abs(a - b)
    d[both_nan | both_inf] = 0.0
    den = np.maximum(np.abs(b), 1e-300)
    e = d / den
    e[(a == b)] = 0.0
    return float(np.max(e)) if e.size else 0.0
