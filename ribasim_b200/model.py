"""Host-side model driver: the role `Ribasim.Model` + the SciML integrator loop play in the
reference (core/src/model.jl:108-297, core/src/bmi.jl:6-99), for an ensemble of members that
share one network.

The time loop is fixed-step implicit Euler.  Per accepted step it reproduces the reference's
callback bookkeeping (core/src/callback.jl:18-72):
RHS caches at the new state -> exact cumulative forcings + tprev -> new vertical fluxes at
forcing tstops.  All floating-point work of the step runs on the GPU through librbs.
"""

from __future__ import annotations

import math
from pathlib import Path

import numpy as np

from . import lib
from .flatten import Flat, flatten
from .network import Params, build_params
from .symbolic import Symbolic, analyse
from .tables import ModelTables
from .timecache import (
    eval_time_params,
    flow_boundary_increment,
    forcing_tstops,
    update_vertical_flux,
)

FORCING_NAMES = ("precipitation", "surface_runoff", "potential_evaporation", "drainage",
                 "infiltration")


class Model:
    """An ensemble of `n_members` instances of one Ribasim network on one GPU."""

    def __init__(self, model: str | Path | ModelTables | Params, n_members: int = 1,
                 device: int = 0, dt: float | None = None, flat: Flat | None = None,
                 sym: Symbolic | None = None, lu_mode: int = -1, lu_tile: int = 0,
                 full_newton: bool = True, predictor: bool = True, max_halvings: int = 20,
                 grow_iters: int = 4, early_exit: bool = True, shrink_log2: int = 2,
                 smem_tile: int = -1, smem_threads: int = 1024,
                 lin_mode: int = -1, n_groups: int = 0, dc_record_cap: int = 64):
        if isinstance(model, Params):
            self.p = model
        else:
            tables = model if isinstance(model, ModelTables) else ModelTables.read(model)
            self.p = build_params(tables)
        self.tables = self.p.tables
        self.flat = flat if flat is not None else flatten(self.p)
        self.sym = sym if sym is not None else analyse(
            self.flat.ints["n_reduced"], self.flat.arrays["wrow_ptr"], self.flat.arrays["wcol"])
        self.B = int(n_members)
        # The device path is fixed-step implicit Euler with Newton sub-stepping; any
        # other `solver.algorithm` (core/src/config.jl:368-380; default QNDF) must be asked for
        # explicitly as ImplicitEuler — nothing is substituted silently.  `dt=` given by the caller
        # is that explicit request.
        alg = self.tables.solver_option("algorithm")
        self.adaptive = dt is None and alg == "QNDF"
        if dt is None and alg not in (None, "ImplicitEuler", "QNDF"):
            raise ValueError(
                f"solver.algorithm = {alg!r} is not available on the device path: QNDF (adaptive, the "
                "reference's default) and fixed-step ImplicitEuler are (set `algorithm` and `dt` in "
                "[solver], or pass dt=)")
        cfg_dt = self.tables.solver_option("dt")
        self.dt = float(dt if dt is not None else (cfg_dt if cfg_dt else 3600.0))
        self.t = 0.0
        self.t_end = self.tables.t_end
        self.tstops = [x for x in self.p.tstops if x <= self.t_end]
        self._ftstops = set(forcing_tstops(self.p, self.t_end))
        self._k_stop = 0
        self.h = lib.Handle(self.flat, self.sym, n_members=self.B, device=device, lu_mode=lu_mode,
                            lu_tile=lu_tile, smem_tile=smem_tile, smem_threads=smem_threads,
                            lin_mode=lin_mode, n_groups=n_groups, dc_rec_cap=dc_record_cap)
        self.DC_RECORD_CAP = int(dc_record_cap)
        self.h.set_solver(self.tables.solver_option("abstol"), self.tables.solver_option("reltol"))
        self.h.set_stepper(full_newton, predictor, max_halvings, grow_iters, early_exit, shrink_log2)
        n = self.p.nodes
        ud = n["user_demand"]
        self.h.set_param("ud_demand_from_timeseries",
                         np.asarray(ud["demand_from_timeseries"], dtype=np.int32))
        alloc_total = np.minimum(ud["demand"], ud["allocated"]).sum(axis=1) \
            if len(ud["node_id"]) else np.zeros(0)
        self.h.set_param("ud_alloc_total", alloc_total)
        self._ud_demand_uploaded = np.array(ud["demand"], dtype=float, copy=True)
        # forcing at t0 is the same for all members until the caller perturbs it
        self._uploaded = {}
        self._overridden: set[str] = set()  # forcing arrays set per member with set_forcing
        for name in FORCING_NAMES:
            self.h.set_member(name, self.p.basin.vertical_flux[name], broadcast=True)
            self._uploaded[name] = self.p.basin.vertical_flux[name].copy()
        self._fb_rate_members: np.ndarray | None = None
        self._push_time_params(0.0)
        self._push_fb_rate(0.0, min(self.dt, self.t_end))
        self.h.set_tprev(0.0)
        # level_prev = level at t0 (core/src/model.jl:187-191)
        self.h.refresh(0.0)
        nb = self.flat.ints["n_basin"]
        self.h.set_member("level_prev", self.h.get_member("level", nb))
        # callbacks are initialised at t = 0: DiscreteControl sets the first control states
        # (core/src/callback.jl:631-634), then the caches are refreshed with them
        if self.flat.ints["n_dc"] > 0:
            self.h.apply_discrete_control()
            # a truth state without a control state is an error at t = 0 already
            # (core/src/callback.jl:671-675); the step path zeroes the status word at every step
            st0 = self.h.get_member_i32("status", 1)[0]
            if np.any(st0 & 4):
                raise RuntimeError("No control state specified for a DiscreteControl truth state at t = 0 "
                                   f"(member {int(np.flatnonzero(st0 & 4)[0])})")
            self.h.refresh(0.0)
        self.status_counts = dict(newton_failed=0, negative_storage=0)
        from .subgrid import Subgrid

        self.subgrid = Subgrid(self.tables, [n.value for n in self.p.basin.node_id])
        self.qndf = None
        if self.adaptive:
            # adaptive QNDF above the seams (qndf.py): per-state reltol vector with the PID
            # integral states unmasked (core/src/util.jl:733-743)
            from .qndf import Qndf

            nst = self.flat.ints["n_state"]
            relmask = np.ones(nst, dtype=bool)
            relmask[self.flat.ints["off_integral"]:] = False
            dtmax = self.tables.solver_option("dtmax")
            self.qndf = Qndf(_SeamBackend(self), np.zeros((nst, self.B)), 0.0,
                             self.tables.solver_option("abstol"), self.tables.solver_option("reltol"),
                             relmask, dtmin=self.tables.solver_option("dtmin") or 0.0,
                             dtmax=dtmax if dtmax else None,
                             maxiters=int(self.tables.solver_option("maxiters")),
                             dt0=cfg_dt if cfg_dt else None)

    # ------------------------------------------------------------------ forcing / parameters
    def _push_time_params(self, t: float) -> None:
        self._tp = eval_time_params(self.p, t)
        self.h.set_time_params(self._tp)

    def _push_fb_rate(self, t0: float, t1: float) -> None:
        nfb = self.flat.ints["n_fb"]
        if nfb == 0:
            return
        if self._fb_rate_members is not None:
            self.h.set_member("fb_rate", self._fb_rate_members)
            return
        incr = flow_boundary_increment(self.p, t0, t1)
        kind = self.p.nodes["flow_boundary"]["flow_rate"][0].kind if nfb else "constant"
        if kind == "constant":
            rate = np.array([s(t0) for s in self.p.nodes["flow_boundary"]["flow_rate"]])
        else:  # linear interpolation: mean rate over the step
            rate = incr / (t1 - t0) if t1 > t0 else incr * 0.0
        self.h.set_member("fb_rate", rate, broadcast=True)

    def set_forcing(self, name: str, values: np.ndarray | None) -> None:
        """Overwrite a Basin vertical flux for all members: values `[n_basin][n_members]`
        (the BMI-writable arrays `basin.infiltration` etc., core/src/bmi.jl:64-88).  Only the named
        array stops following its Basin / time table and the BMI host array; `None` hands it back."""
        if name not in FORCING_NAMES:
            raise KeyError(name)
        if values is None:  # hand the array back to the Basin tables / BMI host arrays
            self._overridden.discard(name)
            self._uploaded[name] = None
            return
        self.h.set_member(name, values)
        self._overridden.add(name)

    def set_member_param(self, name: str, values: np.ndarray | None) -> None:
        """Parameter ensemble: per-member values `[n_nodes][n_members]` of a shared link parameter
        (`mn_manning_n`, `lr_resistance`, `pump_flow_rate`, `outlet_flow_rate`, `lb_level`); None restores the
        shared value of the model tables."""
        self.h.set_member_param(name, values)

    def set_flow_boundary_rate(self, values: np.ndarray | None) -> None:
        """Per-member FlowBoundary rates `[n_fb][n_members]` for the coming steps."""
        self._fb_rate_members = None if values is None else np.ascontiguousarray(values)
        if values is not None:
            self.h.set_member("fb_rate", self._fb_rate_members)

    # ------------------------------------------------------------------ time stepping
    def _next_stop(self, limit: float) -> float:
        while self._k_stop < len(self.tstops) and self.tstops[self._k_stop] <= self.t:
            self._k_stop += 1
        nxt = self.tstops[self._k_stop] if self._k_stop < len(self.tstops) else self.t_end
        return min(nxt, limit)

    def _next_time(self, limit: float) -> float:
        return min(self.t + self.dt, self._next_stop(limit))

    def update(self, want_status: bool = True):
        """One solver step (BMI.update, core/src/bmi.jl:36-39)."""
        self._sync_inputs()
        if self.adaptive:
            return self._qndf_step(self._next_stop(self.t_end))
        return self._step_to(self._next_time(self.t_end), want_status)

    def _qndf_step(self, t_limit: float):
        """One accepted adaptive step that does not pass `t_limit`, then the reference's
        step-boundary callbacks: exact forcing bookkeeping, DiscreteControl, tolerance decrease,
        forcing update at tstops (core/src/callback.jl:7-83)."""
        q = self.qndf
        h_done = q.step(t_limit)
        self.t = q.t
        self.h.accept_step(h_done)
        self.h.set_member("u", q.u)
        self.h.refresh(self.t)  # caches and flows at the new state (results, BMI, DiscreteControl)
        if self.flat.ints["n_dc"] > 0:
            self.h.apply_discrete_control()
        q.decrease_tolerance()
        self._after_step()
        return None

    def _sync_inputs(self) -> None:
        """Upload the BMI-writable host arrays the caller changed since the last step
        (`basin.infiltration` etc. are written through `get_value_ptr` in the reference,
        core/src/bmi.jl:64-88; here the host arrays are mirrors synced at update boundaries)."""
        # `user_demand.demand` (static demands; core/src/bmi.jl:82-83): the flow a UserDemand asks
        # for is min(allocated, demand) summed over the priorities (solve.jl:348-404)
        ud = self.p.nodes["user_demand"]
        if len(ud["node_id"]) and not np.array_equal(ud["demand"], self._ud_demand_uploaded):
            self.h.set_param("ud_alloc_total", np.minimum(ud["demand"], ud["allocated"]).sum(axis=1))
            self._ud_demand_uploaded = np.array(ud["demand"], dtype=float, copy=True)
        for name in FORCING_NAMES:
            if name in self._overridden:
                continue
            arr = self.p.basin.vertical_flux[name]
            if self._uploaded[name] is None or not np.array_equal(arr, self._uploaded[name], equal_nan=True):
                self.h.set_member(name, arr, broadcast=True)
                self._uploaded[name] = arr.copy()

    def _count_status(self, st) -> None:
        if st is not None:
            self.status_counts["newton_failed"] += int(np.count_nonzero(st & 1))
            self.status_counts["negative_storage"] += int(np.count_nonzero(st & 2))
            if np.any(st & 4):
                # core/src/callback.jl:671-675
                m = int(np.flatnonzero(st & 4)[0])
                raise RuntimeError(
                    f"No control state specified for a DiscreteControl truth state (member {m}, t = {self.t})")

    def _after_step(self) -> None:
        """`update_basin!` at forcing tstops (core/src/callback.jl:846-866)."""
        if self.t in self._ftstops:
            update_vertical_flux(self.p, self.t)
            self._sync_inputs()

    def _step_to(self, t_next: float, want_status: bool = True):
        t0 = self.t
        self._push_time_params(t_next)
        self._push_fb_rate(t0, t_next)
        st = self.h.step(t0, t_next - t0, want_status)
        self.t = t_next
        self._count_status(st)
        self._after_step()
        return st

    def _params_constant(self, t_a: float, t_b: float) -> bool:
        """True when the shared time-dependent parameters are the same at both times (block
        interpolation between tstops), so the steps in between can run as one device window."""
        a, b = eval_time_params(self.p, t_a), eval_time_params(self.p, t_b)
        return all(np.array_equal(a[k], b[k], equal_nan=True) for k in a)

    def _fb_rate_constant(self) -> bool:
        """False when FlowBoundary rates are interpolated linearly: every step then gets its own
        exact mean rate (`formulate_flow_boundary!` integrates per step, solve.jl:86-99) instead
        of sharing one rate over a window."""
        if self.flat.ints["n_fb"] == 0 or self._fb_rate_members is not None:
            return True
        return all(s.kind == "constant" for s in self.p.nodes["flow_boundary"]["flow_rate"])

    def update_until(self, time: float, want_status: bool = True) -> None:
        """BMI.update_until (core/src/bmi.jl:41-52).  Whole `dt` steps that end before the next
        tstop run as one `rbs_advance` window; the step that lands on a tstop (or on `time`)
        is taken on its own with the parameters of that instant, like the reference's RHS."""
        if time < self.t:
            raise ValueError("The model has already passed the given timestamp.")
        while self.adaptive and self.t < time:
            self._sync_inputs()
            self._qndf_step(self._next_stop(time))
        while self.t < time:
            self._sync_inputs()
            t_stop = self._next_stop(time)
            n_full = int(math.floor((t_stop - self.t) / self.dt + 1e-9))
            if n_full >= 1 and self.t + n_full * self.dt >= t_stop - 1e-9 * self.dt:
                n_full -= 1  # the step that ends on the stop is taken separately
            if n_full >= 2 and self._fb_rate_constant() and \
                    self._params_constant(self.t + self.dt, self.t + n_full * self.dt):
                t0 = self.t
                self._push_time_params(t0 + self.dt)
                self._push_fb_rate(t0, t0 + n_full * self.dt)
                st = self.h.advance(t0, self.dt, n_full, want_status)
                self.t = t0 + n_full * self.dt
                self._count_status(st)
                self._after_step()
            else:
                self._step_to(self._next_time(time), want_status)

    def run(self, saveat: float | None = None, results_dir: str | Path | None = None,
            check_water_balance: bool = True):
        """`Ribasim.run` for the ensemble (core/src/main.jl, model.jl:285-297).  With `saveat`
        (seconds; default the TOML's `solver.saveat`) or `results_dir` the run records the
        reference's per-period outputs (`results.Results`: storages, levels, mean flows and the
        water balance per Basin, callback.jl:390-558) and, given a directory, writes basin.nc,
        flow.nc and basin_state.nc (write.jl:6-67); returns the recorder."""
        if saveat is None and results_dir is None:
            self.update_until(self.t_end)
            return None
        from .results import Results

        step = float(saveat if saveat is not None else self.tables.solver_option("saveat"))
        if not step > 0:  # saveat = 0: every solver step
            step = self.dt
        res = Results(self, check_water_balance)
        t0, k = self.t, 0
        while self.t < self.t_end:
            k += 1
            self.update_until(min(t0 + k * step, self.t_end))
            res.save()
        if results_dir is not None:
            res.write(results_dir)
        return res

    # ------------------------------------------------------------------ results
    def storage(self) -> np.ndarray:
        return self.h.get_member("storage", self.flat.ints["n_basin"])

    def level(self) -> np.ndarray:
        return self.h.get_member("level", self.flat.ints["n_basin"])

    def u(self) -> np.ndarray:
        return self.h.get_member("u", self.flat.ints["n_state"])

    def subgrid_level(self) -> np.ndarray:
        """`update_subgrid_level!` (core/src/callback.jl:788-813): `[n_subgrid][n_members]` from
        the current Basin levels and the h(h) relations in force at the current time."""
        return self.subgrid.levels(self.level(), self.t)

    def control_state(self) -> np.ndarray:
        """Control state index per DiscreteControl node and member (-1 = none); the names are
        `flat.dc_state_names[i]`."""
        return self.h.get_member_i32("dc_state", self.flat.ints["n_dc"])

    DC_RECORD_CAP = 64  # control state changes recorded per member (constructor: dc_record_cap)

    def control_record(self, member: int = 0) -> list[tuple[float, int, str, str]]:
        """`discrete_control.record` of one member (core/src/callback.jl:698-716):
        `(time, control_node_id, truth_state, control_state)` per control state change, the
        initial states at t = 0 included, in time order (ties in node order)."""
        n_dc = self.flat.ints["n_dc"]
        if n_dc == 0:
            return []
        cap = self.DC_RECORD_CAP
        cnt = int(self.h.get_member_i32("cnt_dc", 1)[0, member])
        if cnt > cap:
            import warnings

            warnings.warn(f"member {member} changed control state {cnt} times; only the first {cap} are recorded "
                          "(Model(..., dc_record_cap=N) keeps more)")
            cnt = cap
        t = self.h.get_member("dc_rec_t", cap)[:cnt, member]
        ns = self.h.get_member_i32("dc_rec_ns", cap)[:cnt, member]
        bits = self.h.get_member_i32("dc_rec_bits", cap)[:cnt, member]
        ids = [x.value for x in self.p.nodes["discrete_control"]["node_id"]]
        n_cond = np.diff(self.flat.arrays["dc_cond_ptr"])
        rec = []
        for k in sorted(range(cnt), key=lambda k: (t[k], int(ns[k]) & 0xFFFF)):
            i, st = int(ns[k]) & 0xFFFF, int(ns[k]) >> 16
            truth = "".join("T" if (int(bits[k]) >> j) & 1 else "F" for j in range(int(n_cond[i])))
            rec.append((float(t[k]), ids[i], truth, self.flat.dc_state_names[i][st]))
        return rec

    def flow(self) -> np.ndarray:
        return self.h.get_member("du", self.flat.ints["n_state"])

    def finalize(self) -> None:
        self.h.close()


class _SeamBackend:
    """The device behind the SciML plug-in seams of the C ABI, as `qndf.Qndf` uses it."""

    def __init__(self, model: "Model"):
        self.m = model
        self.h = model.h
        f = model.flat
        self.jcol = np.asarray(f.arrays["jcol"], dtype=np.int64)
        rp = np.asarray(f.arrays["jrow_ptr"], dtype=np.int64)
        self.jrow = np.repeat(np.arange(f.ints["n_state"]), np.diff(rp))
        self.n = f.ints["n_state"]
        self.jv = None

    def prepare(self, t0: float, t1: float) -> None:
        self.m._push_time_params(t1)
        self.m._push_fb_rate(t0, t1)

    def reduce(self, u):
        return self.h.reduce(u)

    def rhs(self, ur, t):
        du, c = self.h.rhs(ur, t, with_cache=True)
        return du, c["storage"]

    def factor(self, ur, t, gamma):
        self.jv, _ = self.h.jac(ur, t, gamma)

    def solve(self, b):
        return self.h.solve(b)

    def jmul(self, c):
        out = np.zeros((self.n, c.shape[1]))
        np.add.at(out, self.jrow, self.jv * c[self.jcol])
        return out

    def limit(self, u, uprev, dt, t):
        return self.h.limit_flow(u, uprev, dt, t)
