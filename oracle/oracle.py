"""ctypes binding of the CPU oracle (oracle/rbs_oracle.cpp).

TEST INFRASTRUCTURE ONLY — imported by tests/, `__graft_entry__.smoke()` and bench.py's CPU
baseline legs.  The product package `ribasim_b200` never imports this module.

The oracle gets its *own*, reference-shaped description of the model (per node type: the
inflow / outflow node of every connector, per Basin the inflow and outflow neighbour lists in
reference order) built here from the host parameters, not the flattened CSR layout that the
CUDA library consumes, so layout bugs in the flattening do not cancel out in parity tests.
"""

from __future__ import annotations

import ctypes
import subprocess
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
LIB_PATH = HERE / "librbs_oracle.so"

_TYPE_CODE = {
    "Basin": 0, "LevelBoundary": 1, "Terminal": 2, "TabulatedRatingCurve": 3, "Pump": 4,
    "Outlet": 5, "UserDemand": 6, "LinearResistance": 7, "ManningResistance": 8,
    "FlowBoundary": 9,
}


def build(force: bool = False) -> Path:
    src = HERE / "rbs_oracle.cpp"
    if force or not LIB_PATH.exists() or LIB_PATH.stat().st_mtime < src.stat().st_mtime:
        subprocess.run(["make", "-C", str(HERE), "-B", "librbs_oracle.so"], check=True,
                       capture_output=True)
    return LIB_PATH


_lib = None


def lib() -> ctypes.CDLL:
    global _lib
    if _lib is None:
        build()
        L = ctypes.CDLL(str(LIB_PATH))
        vp, cp, dp, ip = ctypes.c_void_p, ctypes.c_char_p, ctypes.POINTER(ctypes.c_double), \
            ctypes.POINTER(ctypes.c_int32)
        L.orc_create.restype = vp
        L.orc_clone.restype = vp
        L.orc_clone.argtypes = [vp]
        L.orc_destroy.argtypes = [vp]
        L.orc_set_f64.argtypes = [vp, cp, dp, ctypes.c_int]
        L.orc_set_i32.argtypes = [vp, cp, ip, ctypes.c_int]
        L.orc_get_f64.argtypes = [vp, cp, dp, ctypes.c_int]
        L.orc_finalize.argtypes = [vp]
        L.orc_set_pattern.argtypes = [vp, ip, ip, ctypes.c_int]
        for f in ("orc_n_colors", "orc_n_state", "orc_n_reduced"):
            getattr(L, f).argtypes = [vp]
        L.orc_set_tprev.argtypes = [vp, ctypes.c_double]
        L.orc_get_tprev.argtypes = [vp]
        L.orc_get_tprev.restype = ctypes.c_double
        L.orc_set_merged_exact.argtypes = [vp, ctypes.c_int]
        L.orc_set_level_prev.argtypes = [vp, dp]
        L.orc_reduce.argtypes = [vp, dp, dp]
        L.orc_rhs.argtypes = [vp, dp, ctypes.c_double, dp, dp, dp, dp, dp]
        L.orc_jac_dense.argtypes = [vp, dp, ctypes.c_double, dp]
        L.orc_jac_colored.argtypes = [vp, dp, ctypes.c_double, dp]
        L.orc_W_inner.argtypes = [vp, dp, ctypes.c_double, dp]
        L.orc_dense_solve.argtypes = [ctypes.c_int, dp, dp, dp]
        L.orc_limit_flow.argtypes = [vp, dp, dp, ctypes.c_double, ctypes.c_double]
        L.orc_step.argtypes = [vp, dp, ctypes.c_double, ctypes.c_double, ctypes.c_double,
                               ctypes.c_double, ctypes.c_int, dp, ip, ip, ip]
        L.orc_apply_discrete_control.argtypes = [vp, dp, ctypes.c_double]
        L.orc_get_dc.argtypes = [vp, ip, ip, ip]
        L.orc_get_dc_record.argtypes = [vp, ip, ctypes.c_int]
        L.orc_advance.argtypes = [vp, dp, ctypes.c_double, ctypes.c_double, dp, dp, dp, ip]
        L.orc_reduction_factor.argtypes = [ctypes.c_double, ctypes.c_double]
        L.orc_reduction_factor.restype = ctypes.c_double
        L.orc_relaxed_root.argtypes = [ctypes.c_double, ctypes.c_double]
        L.orc_relaxed_root.restype = ctypes.c_double
        L.orc_storage_to_level.argtypes = [vp, ctypes.c_int, ctypes.c_double]
        L.orc_storage_to_level.restype = ctypes.c_double
        L.orc_level_to_area.argtypes = [vp, ctypes.c_int, ctypes.c_double]
        L.orc_level_to_area.restype = ctypes.c_double
        L.orc_pchip_eval.argtypes = [vp, ctypes.c_int, ctypes.c_double]
        L.orc_pchip_eval.restype = ctypes.c_double
        L.orc_run_batch.argtypes = [ctypes.POINTER(vp), ctypes.POINTER(dp), ctypes.c_int,
                                    ctypes.c_double, ctypes.c_double, ctypes.c_int, dp, dp, dp,
                                    ctypes.c_int, ctypes.POINTER(ctypes.c_int64)]
        _lib = L
    return _lib


def _dp(a: np.ndarray):
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_double))


def _ip(a: np.ndarray):
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_int32))


def _f64(x) -> np.ndarray:
    return np.ascontiguousarray(np.asarray(x, dtype=np.float64)).reshape(-1)


def _i32(x) -> np.ndarray:
    return np.ascontiguousarray(np.asarray(x, dtype=np.int32)).reshape(-1)


class OracleModel:
    """One model instance (= one ensemble member) of the CPU restatement."""

    def __init__(self, p, handle=None):
        self.L = lib()
        self.p = p
        if handle is not None:
            self.h = handle
            self._own = True
            self.n_state = self.L.orc_n_state(self.h)
            self.n_reduced = self.L.orc_n_reduced(self.h)
            return
        self.h = self.L.orc_create()
        self._own = True
        self._describe(p)
        self.L.orc_finalize(self.h)
        self.n_state = self.L.orc_n_state(self.h)
        self.n_reduced = self.L.orc_n_reduced(self.h)
        assert self.n_state == p.n_states, (self.n_state, p.n_states)

    def clone(self) -> "OracleModel":
        return OracleModel(self.p, handle=self.L.orc_clone(self.h))

    def __del__(self):
        try:
            if getattr(self, "_own", False) and self.h:
                self.L.orc_destroy(self.h)
                self.h = None
        except Exception:
            pass

    # ------------------------------------------------------------------ description
    def set_f64(self, key: str, arr) -> None:
        a = _f64(arr)
        self.L.orc_set_f64(self.h, key.encode(), _dp(a), a.size)

    def set_i32(self, key: str, arr) -> None:
        a = _i32(arr)
        self.L.orc_set_i32(self.h, key.encode(), _ip(a), a.size)

    def get_f64(self, key: str, n: int) -> np.ndarray:
        a = np.zeros(n)
        rc = self.L.orc_get_f64(self.h, key.encode(), _dp(a), n)
        if rc:
            raise KeyError(key)
        return a

    def _describe(self, p) -> None:
        n = p.nodes
        b = p.basin
        code = lambda nid: (_TYPE_CODE.get(nid.type, 10), nid.idx - 1)  # noqa: E731
        ud = n["user_demand"]
        counts = dict(
            n_basin=len(b.node_id), n_trc=len(n["tabulated_rating_curve"]["node_id"]),
            n_pump=len(n["pump"]["node_id"]), n_outlet=len(n["outlet"]["node_id"]),
            n_ud=len(ud["node_id"]), n_lr=len(n["linear_resistance"]["node_id"]),
            n_manning=len(n["manning_resistance"]["node_id"]),
            n_pid=len(n["pid_control"]["node_id"]), n_cc=len(n["continuous_control"]["node_id"]),
            n_lb=len(n["level_boundary"]["node_id"]), n_fb=len(n["flow_boundary"]["node_id"]),
            n_prio=len(ud["demand_priorities"]),
        )
        for k, v in counts.items():
            self.set_i32(k, [v])
        self.set_f64("level_difference_threshold", [p.level_difference_threshold])
        # basins
        self.set_f64("storage0", b.storage0)
        self.set_f64("low_storage_threshold", b.low_storage_threshold)
        self.set_f64("fixed_area", [a[-1] for a in b.areas])
        ptr = np.cumsum([0] + [len(l) for l in b.levels])
        self.set_i32("prof_ptr", ptr)
        self.set_f64("prof_level", np.concatenate(b.levels))
        self.set_f64("prof_area", np.concatenate(b.areas))
        self.set_f64("prof_dSdh", np.concatenate(b.dSdh))
        self.set_f64("prof_cumS", np.concatenate(b.cumS))
        for name, lists in (("bin", b.inflow_ids), ("bout", b.outflow_ids)):
            self.set_i32(f"{name}_ptr", np.cumsum([0] + [len(l) for l in lists]))
            flat = [code(x) for l in lists for x in l]
            self.set_i32(f"{name}_type", [c[0] for c in flat])
            self.set_i32(f"{name}_idx", [c[1] for c in flat])
        for name, v in b.vertical_flux.items():
            self.set_f64(name, v)
        # connectors
        for key, pre in (("tabulated_rating_curve", "trc"), ("pump", "pump"), ("outlet", "outlet"),
                         ("linear_resistance", "lr"), ("manning_resistance", "mn")):
            d = n[key]
            src = [code(l.link[0]) for l in d["inflow_link"]]
            dst = [code(l.link[1]) for l in d["outflow_link"]]
            self.set_i32(f"{pre}_src_type", [c[0] for c in src])
            self.set_i32(f"{pre}_src_idx", [c[1] for c in src])
            self.set_i32(f"{pre}_dst_type", [c[0] for c in dst])
            self.set_i32(f"{pre}_dst_idx", [c[1] for c in dst])
        self.set_f64("lr_resistance", n["linear_resistance"]["resistance"])
        self.set_f64("lr_max_flow_rate", n["linear_resistance"]["max_flow_rate"])
        mn = n["manning_resistance"]
        for k in ("length", "manning_n", "profile_width", "profile_slope", "upstream_bottom",
                  "downstream_bottom"):
            self.set_f64(f"mn_{k}", mn[k])
        # PCHIP tables: rating curves first, then continuous-control functions
        trc = n["tabulated_rating_curve"]
        cc = n["continuous_control"]
        tabs = list(trc["interpolations"]) + list(cc["func"])
        self.set_i32("tab_ptr", np.cumsum([0] + [len(t.t) for t in tabs]))
        self.set_f64("tab_t", np.concatenate([t.t for t in tabs]) if tabs else [])
        self.set_f64("tab_u", np.concatenate([t.u for t in tabs]) if tabs else [])
        self.set_i32("tab_left_linear", [t.left == "linear" for t in tabs])
        self.set_i32("tab_right_linear", [t.right == "linear" for t in tabs])
        for key in ("pump", "outlet"):
            self.set_i32(f"{key}_control_type", n[key]["control_type"])
        # user demand
        links = [l for ls in ud["inflow_links"] for l in ls]
        self.set_i32("ud_link_ptr", ud["inflow_link_offsets"])
        self.set_i32("ud_link_src_type", [code(l.link[0])[0] for l in links])
        self.set_i32("ud_link_src_idx", [code(l.link[0])[1] for l in links])
        self.set_i32("ud_dst_type", [code(l.link[1])[0] for l in ud["outflow_link"]])
        self.set_i32("ud_dst_idx", [code(l.link[1])[1] for l in ud["outflow_link"]])
        self.set_f64("ud_min_level", ud["min_level"])
        self.set_f64("ud_link_alloc", [x for l in ud["inflow_link_allocated"] for x in l])
        self.set_i32("ud_has_priority", np.asarray(ud["has_demand_priority"]).astype(np.int32))
        self.set_f64("ud_allocated", ud["allocated"])
        self.set_f64("ud_demand_static", ud["demand"])
        self.set_i32("ud_demand_from_timeseries", ud["demand_from_timeseries"])
        # flow boundaries
        fb = n["flow_boundary"]
        self.set_i32("fb_dst_type", [code(l.link[1])[0] for l in fb["outflow_link"]])
        self.set_i32("fb_dst_idx", [code(l.link[1])[1] for l in fb["outflow_link"]])
        # PID
        pid = n["pid_control"]
        self.set_i32("pid_listen_basin", [l.idx - 1 for l in pid["listen_node_id"]])
        self.set_i32("pid_target_is_outlet", [t.type == "Outlet" for t in pid["target_node"]])
        self.set_i32("pid_target_idx", [t.idx - 1 for t in pid["target_node"]])
        # continuous control
        self.set_i32("cc_sv_ptr", np.cumsum([0] + [len(c) for c in cc["compound_variable"]]))
        svs = [sv for c in cc["compound_variable"] for sv in c]
        self.set_i32("cc_sv_kind", [sv["kind"] for sv in svs])
        self.set_i32("cc_sv_idx", [sv["idx"] for sv in svs])
        self.set_f64("cc_sv_weight", [sv["weight"] for sv in svs])
        self.set_i32("cc_table_idx", [len(trc["interpolations"]) + i for i in range(len(cc["func"]))])
        self.set_i32("cc_target_is_outlet", [t.type == "Outlet" for t in cc["target_node"]])
        self.set_i32("cc_target_idx", [t.idx - 1 for t in cc["target_node"]])
        # DiscreteControl: shared description (index layout of flatten._flatten_discrete_control)
        from ribasim_b200.flatten import flatten
        flat = flatten(p)
        self.set_i32("n_dc", [flat.ints["n_dc"]])
        for k, v in flat.arrays.items():
            if k.startswith("dc_") or "_cs_" in k or k.endswith("_dc"):
                (self.set_i32 if v.dtype == np.int32 else self.set_f64)(k, v)
        self.n_dc = flat.ints["n_dc"]
        self.n_dc_cond = flat.ints["n_dc_cond"]

    # ------------------------------------------------------------------ time cache
    def set_time_params(self, tp: dict[str, np.ndarray]) -> None:
        for k, v in tp.items():
            if k in ("trc_table_idx",):
                self.set_i32(k, v)
            elif k == "ud_total_demand":
                continue
            else:
                self.set_f64(k, v)

    def apply_discrete_control(self, u_reduced, t: float) -> None:
        ur = _f64(u_reduced)
        self.L.orc_apply_discrete_control(self.h, _dp(ur), t)

    def discrete_control_state(self):
        st = np.zeros(max(self.n_dc, 1), dtype=np.int32)
        tr = np.zeros(max(self.n_dc_cond, 1), dtype=np.int32)
        cnt = np.zeros(2, dtype=np.int32)
        self.L.orc_get_dc(self.h, _ip(st), _ip(tr), _ip(cnt))
        return st[: self.n_dc], tr[: self.n_dc_cond], int(cnt[0]), int(cnt[1])

    def discrete_control_record(self) -> list[tuple[int, int, int]]:
        """(DiscreteControl index, truth bits, new control state index) per change."""
        buf = np.zeros(3 * 4096, dtype=np.int32)
        n = self.L.orc_get_dc_record(self.h, _ip(buf), buf.size)
        return [tuple(int(x) for x in buf[i:i + 3]) for i in range(0, min(n, buf.size), 3)]

    def set_pattern(self, rows_1based, cols_1based) -> None:
        r = _i32(np.asarray(rows_1based) - 1)
        c = _i32(np.asarray(cols_1based) - 1)
        self.L.orc_set_pattern(self.h, _ip(r), _ip(c), r.size)
        self.nnz = r.size

    # ------------------------------------------------------------------ operations
    def reduce(self, u) -> np.ndarray:
        u = _f64(u)
        out = np.zeros(self.n_reduced)
        self.L.orc_reduce(self.h, _dp(u), _dp(out))
        return out

    def rhs(self, u_reduced, t: float, with_cache: bool = False):
        ur = _f64(u_reduced)
        du = np.zeros(self.n_state)
        nb = len(self.p.basin.node_id)
        st, lv, ar, fc = (np.zeros(nb) for _ in range(4))
        self.L.orc_rhs(self.h, _dp(ur), t, _dp(du), _dp(st), _dp(lv), _dp(ar), _dp(fc))
        if with_cache:
            return du, dict(storage=st, level=lv, area=ar, factor=fc)
        return du

    def jac_dense(self, u_reduced, t: float) -> np.ndarray:
        ur = _f64(u_reduced)
        J = np.zeros((self.n_reduced, self.n_state))
        self.L.orc_jac_dense(self.h, _dp(ur), t, _dp(J))
        return J.T  # (n_state, n_reduced)

    def jac_colored(self, u_reduced, t: float) -> np.ndarray:
        ur = _f64(u_reduced)
        vals = np.zeros(self.nnz)
        self.L.orc_jac_colored(self.h, _dp(ur), t, _dp(vals))
        return vals

    def W_inner(self, vals, gamma: float) -> np.ndarray:
        v = _f64(vals)
        W = np.zeros((self.n_reduced, self.n_reduced))
        self.L.orc_W_inner(self.h, _dp(v), gamma, _dp(W))
        return W.T  # column-major -> [i, j]

    def dense_solve(self, W: np.ndarray, b) -> np.ndarray:
        n = W.shape[0]
        Wc = np.ascontiguousarray(W.T, dtype=np.float64)  # column-major
        b = _f64(b)
        x = np.zeros(n)
        rc = self.L.orc_dense_solve(n, _dp(Wc.reshape(-1)), _dp(b), _dp(x))
        if rc:
            raise np.linalg.LinAlgError("singular W")
        return x

    def limit_flow(self, u, uprev, dt: float, t: float) -> np.ndarray:
        u = _f64(u).copy()
        up = _f64(uprev)
        self.L.orc_limit_flow(self.h, _dp(u), _dp(up), dt, t)
        return u

    def step(self, u: np.ndarray, t: float, dt: float, abstol: float, reltol: float,
             eta_old: float = 1.0, max_iter: int = 10):
        """One implicit-Euler step in place on `u`; returns (rc, iters, converged, neg, eta)."""
        assert u.dtype == np.float64 and u.flags.c_contiguous
        eta = ctypes.c_double(eta_old)
        it, cv, ng = ctypes.c_int32(0), ctypes.c_int32(0), ctypes.c_int32(0)
        rc = self.L.orc_step(self.h, _dp(u), t, dt, abstol, reltol, max_iter,
                             ctypes.byref(eta), ctypes.byref(it), ctypes.byref(cv), ctypes.byref(ng))
        return rc, it.value, cv.value, ng.value, eta.value

    def advance(self, u: np.ndarray, t0: float, dt: float, opts: np.ndarray, h_try: float,
                eta_old: float = 1.0):
        """One outer step with sub-stepping, in place on `u`.
        Returns (rc, dict(iters, substeps, rejects, negative, converged), h_try, eta)."""
        assert u.dtype == np.float64 and u.flags.c_contiguous
        h = ctypes.c_double(h_try)
        eta = ctypes.c_double(eta_old)
        out = np.zeros(5, dtype=np.int32)
        rc = self.L.orc_advance(self.h, _dp(u), t0, dt, _dp(opts), ctypes.byref(h),
                                ctypes.byref(eta), _ip(out))
        return rc, dict(iters=int(out[0]), substeps=int(out[1]), rejects=int(out[2]),
                        negative=int(out[3]), converged=int(out[4])), h.value, eta.value


def advance_opts(abstol=1e-5, reltol=1e-5, max_iter=10, full_newton=0, grow_iters=4,
                 max_halvings=0, predictor=0, early_exit=0, shrink_log2=1) -> np.ndarray:
    return np.array([abstol, reltol, max_iter, full_newton, grow_iters, max_halvings, predictor,
                     early_exit, shrink_log2], dtype=np.float64)


class OracleBatch:
    """Independent oracle instances (one per ensemble member) advanced on host threads: the
    CPU baseline of bench.py."""

    def __init__(self, members: list[OracleModel], opts: np.ndarray, n_threads: int):
        self.members = members
        self.opts = np.ascontiguousarray(opts, dtype=np.float64)
        self.n_threads = int(n_threads)
        n = len(members)
        self.us = [np.zeros(m.n_state) for m in members]
        self.h_try = np.zeros(n)
        self.eta = np.ones(n)
        self._hs = (ctypes.c_void_p * n)(*[m.h for m in members])
        self._us = (ctypes.POINTER(ctypes.c_double) * n)(*[_dp(u) for u in self.us])
        self.L = lib()

    def run(self, t0: float, dt: float, n_steps: int) -> dict[str, int]:
        if not np.any(self.h_try):
            self.h_try[:] = dt
        out = (ctypes.c_int64 * 3)()
        self.L.orc_run_batch(self._hs, self._us, len(self.members), t0, dt, n_steps,
                             _dp(self.opts), _dp(self.h_try), _dp(self.eta), self.n_threads, out)
        return dict(failed=out[0], substeps=out[1], iters=out[2])


def run_oracle(p, dt: float, t_end: float | None = None, flat=None, abstol: float | None = None,
               reltol: float | None = None, merged_exact: bool = False, max_steps: int | None = None,
               full_newton: int = 0, max_halvings: int = 0, grow_iters: int = 4, predictor: int = 0):
    """Fixed-step implicit-Euler run of one instance with the reference's step bookkeeping
    (SURVEY Appendix A.6).  Returns (u, storage, level, stats)."""
    from ribasim_b200.flatten import flatten
    from ribasim_b200.timecache import (eval_time_params, flow_boundary_increment,
                                        forcing_tstops, update_vertical_flux)

    if flat is None:
        flat = flatten(p)
    m = p.tables
    t_end = m.t_end if t_end is None else t_end
    abstol = m.solver_option("abstol") if abstol is None else abstol
    reltol = m.solver_option("reltol") if reltol is None else reltol
    om = OracleModel(p)
    om.set_pattern(flat.jac_rows_csc, flat.jac_cols_csc)
    om.L.orc_set_merged_exact(om.h, int(merged_exact))
    u = np.zeros(om.n_state)
    om.set_time_params(eval_time_params(p, 0.0))
    _, cache = om.rhs(om.reduce(u), 0.0, with_cache=True)
    lv = np.ascontiguousarray(cache["level"])
    om.L.orc_set_level_prev(om.h, _dp(lv))
    om.apply_discrete_control(om.reduce(u), 0.0)  # callbacks are initialised at t = 0
    tstops = [x for x in p.tstops if x <= t_end]
    ftstops = set(forcing_tstops(p, t_end))
    t = 0.0
    k_stop = 0
    eta = 1.0
    stats = dict(steps=0, iters=0, failed=0, negative=0, substeps=0, rejects=0)
    opts = advance_opts(abstol, reltol, 10, full_newton, grow_iters, max_halvings, predictor)
    h_try = dt
    while t < t_end:
        while k_stop < len(tstops) and tstops[k_stop] <= t:
            k_stop += 1
        t_next = min(t + dt, tstops[k_stop] if k_stop < len(tstops) else t_end, t_end)
        om.set_time_params(eval_time_params(p, t_next))
        om.set_f64("fb_incr", flow_boundary_increment(p, t, t_next))
        rc, st, h_try, eta = om.advance(u, t, t_next - t, opts, min(h_try, t_next - t), eta)
        stats["steps"] += 1
        stats["iters"] += st["iters"]
        stats["failed"] += int(rc != 0)
        stats["negative"] += st["negative"]
        stats["substeps"] += st["substeps"]
        stats["rejects"] += st["rejects"]
        t = t_next
        if t in ftstops and update_vertical_flux(p, t):
            for name, v in p.basin.vertical_flux.items():
                om.set_f64(name, v)
        if max_steps is not None and stats["steps"] >= max_steps:
            break
    om.set_time_params(eval_time_params(p, t))
    om.set_f64("fb_incr", np.zeros(len(p.nodes["flow_boundary"]["node_id"])))
    _, cache = om.rhs(om.reduce(u), t, with_cache=True)
    stats["oracle"] = om
    return u, cache["storage"], cache["level"], stats
