"""Results of an ensemble run in the reference's output layout: the `save_flow` bookkeeping per
`saveat` period (core/src/callback.jl:390-558), the `basin_data` / `flow_data` / `basin_state_data`
tables (core/src/write.jl:329-532) and `basin.nc`, `flow.nc`, `basin_state.nc` files with the
reference's variable names, dimensions and CF attributes (write.jl:69-83, 159-327, 819-844).

Everything here is host-side arithmetic on what the device path hands back at save times (the
cumulative-flow state `u`, storages, levels and the exactly integrated forcing volumes); the
member axis is kept, so one recorder serves the whole ensemble.  The files are NetCDF classic
(scipy); an ensemble gets a leading `realization` dimension.
"""

from __future__ import annotations

from datetime import timedelta
from pathlib import Path

import numpy as np

from .flatten import KIND_BASIN

RESULTS_FILENAME = dict(basin_state="basin_state.nc", basin="basin.nc", flow="flow.nc",
                        subgrid_level="subgrid_level.nc", control="control.nc")

CF_GLOBAL = {"Conventions": "CF-1.12", "references": "https://ribasim.org"}
_RATE = "m3 s-1"
CF = {
    "node_id": dict(cf_role="timeseries_id", long_name="node identifier"),
    "link_id": dict(cf_role="timeseries_id", long_name="link identifier"),
    "from_node_id": dict(long_name="identifier of the node the link starts at"),
    "to_node_id": dict(long_name="identifier of the node the link ends at"),
    "time": dict(axis="T", calendar="standard", standard_name="time", long_name="time"),
    "level": dict(units="m", standard_name="water_surface_height_above_reference_datum",
                  long_name="water level above reference datum"),
    "storage": dict(units="m3", standard_name="surface_water_amount", long_name="water storage volume"),
    "flow_rate": dict(units=_RATE, standard_name="water_volume_transport_in_river_channel",
                      long_name="water flow rate"),
    "inflow_rate": dict(units=_RATE, standard_name="water_volume_transport_in_river_channel",
                        long_name="water inflow rate"),
    "outflow_rate": dict(units=_RATE, standard_name="water_volume_transport_in_river_channel",
                         long_name="water outflow rate"),
    "storage_rate": dict(units=_RATE, long_name="water storage rate of change"),
    "precipitation": dict(units=_RATE, standard_name="lwe_precipitation_rate",
                          long_name="liquid water equivalent precipitation rate"),
    "surface_runoff": dict(units=_RATE, standard_name="surface_runoff_flux", long_name="surface runoff rate"),
    "evaporation": dict(units=_RATE, standard_name="water_evaporation_flux", long_name="evaporation rate"),
    "drainage": dict(units=_RATE, long_name="drainage rate"),
    "infiltration": dict(units=_RATE, long_name="infiltration rate"),
    "subgrid_id": dict(cf_role="timeseries_id", long_name="subgrid element identifier"),
    "subgrid_level": dict(units="m", standard_name="water_surface_height_above_reference_datum",
                          long_name="subgrid water level above reference datum"),
    "balance_error": dict(units=_RATE, long_name="water balance error"),
    "relative_error": dict(units="1", long_name="relative water balance error"),
}

BASIN_VARS = ("level", "storage", "inflow_rate", "outflow_rate", "storage_rate", "precipitation",
              "surface_runoff", "evaporation", "drainage", "infiltration", "balance_error",
              "relative_error")


class WaterBalanceError(RuntimeError):
    pass


def _node_state(p, nid, inflow: bool = True) -> int | None:
    """`get_state_index(state_ranges, id; inflow)` (core/src/util.jl:878-895), 0-based."""
    from .tables import snake_case

    name = ("user_demand_inflow" if inflow else "user_demand_outflow") if nid.type == "UserDemand" \
        else snake_case(nid.type)
    if name not in p.state_ranges:
        return None
    start, stop = p.state_ranges[name]
    return start + nid.idx - 1 if start + nid.idx - 1 < stop else None


def _state_index(p, link) -> int:
    """`get_state_index(state_ranges, link_to_state_idx, link)` (core/src/util.jl:904-913)."""
    idx = p.link_to_state_idx.get(link)
    if idx is None:
        idx = _node_state(p, link[1])
    if idx is None:
        idx = _node_state(p, link[0], inflow=False)
    if idx is None:
        raise KeyError(f"no state carries the flow of link {link[0]} -> {link[1]}")
    return idx


class Results:
    """Recorder of one `Model`: `save()` at every `saveat` boundary, then the tables / files.

    Arrays are `[n_periods][entity][n_members]`; a period's storages and levels are those at its
    START and its rates the means over the period, as in `basin_data` (write.jl:368-435)."""

    def __init__(self, model, check_water_balance: bool = True):
        self.m = model
        p, flat = model.p, model.flat
        self.nb, self.ns = flat.ints["n_basin"], flat.ints["n_state"]
        self.n_conn, self.n_fb = flat.ints["n_conn"], flat.ints["n_fb"]
        A = flat.arrays
        self._src = np.where(A["src_kind"][: self.n_conn] == KIND_BASIN, A["src_idx"][: self.n_conn], -1)
        self._dst = np.where(A["dst_kind"][: self.n_conn] == KIND_BASIN, A["dst_idx"][: self.n_conn], -1)
        self._fb_basin = np.asarray(A["fb_basin"], dtype=np.int64)
        r = p.state_ranges
        self._evap = slice(*r["evaporation"])
        self._infil = slice(*r["infiltration"])
        self.check = check_water_balance
        self.abstol = float(p.tables.solver_option("water_balance_abstol") or 1e-3)
        self.reltol = float(p.tables.solver_option("water_balance_reltol") or 1e-2)
        # link bookkeeping of flow_data (write.jl:455-532)
        ext, internal = p.external_flow_links, p.internal_flow_links
        ext_pos = {l.id: k for k, l in enumerate(ext)}
        int_pos = {l.id: k for k, l in enumerate(internal)}
        self.link_id = np.array([l.id for l in ext], dtype=np.int32)
        self.from_node_id = np.array([l.link[0].value for l in ext], dtype=np.int32)
        self.to_node_id = np.array([l.link[1].value for l in ext], dtype=np.int32)
        self._map = np.zeros((len(ext), len(internal)))
        for e_id, i_id in p.graph.flow_link_pairs:
            self._map[ext_pos[e_id], int_pos[i_id]] = 1.0
        # internal link -> (+state index) or (-(flow boundary index) - 1)
        self._int_src = []
        for l in internal:
            a = l.link[0]
            self._int_src.append(-a.idx if a.type == "FlowBoundary" else _state_index(p, l.link) + 1)
        self._int_src = np.array(self._int_src, dtype=np.int64)
        self.node_id = np.array([n.value for n in p.basin.node_id], dtype=np.int32)
        self.times = [float(model.t)]
        self.periods: list[dict[str, np.ndarray]] = []
        self._prev = self._snapshot()
        # save_subgrid_level: at the start and at every saveat (callback.jl:60-64, 816-822)
        self.subgrid_levels = [model.subgrid.levels(self._prev["level"], self.times[0])]

    # ------------------------------------------------------------------ bookkeeping
    def _snapshot(self) -> dict[str, np.ndarray]:
        h, B = self.m.h, self.m.B
        snap = dict(u=h.get_member("u", self.ns), storage=h.get_member("storage", self.nb),
                    level=h.get_member("level", self.nb),
                    cp=h.get_member("cum_precipitation", self.nb),
                    cr=h.get_member("cum_surface_runoff", self.nb),
                    cd=h.get_member("cum_drainage", self.nb))
        snap["fb"] = h.get_member("fb_cum", self.n_fb) if self.n_fb else np.zeros((0, B))
        return snap

    def save(self) -> dict[str, np.ndarray]:
        """`save_flow` + `check_water_balance_error!` over the period since the last save."""
        new, old = self._snapshot(), self._prev
        t = float(self.m.t)
        dt = t - self.times[-1]
        if not dt > 0:
            raise ValueError("Results.save: no time has passed since the last save")
        B = self.m.B
        flow = (new["u"] - old["u"]) / dt
        inflow, outflow = np.zeros((self.nb, B)), np.zeros((self.nb, B))
        q = flow[: self.n_conn]
        pos, neg = np.maximum(q, 0.0), np.maximum(-q, 0.0)
        s = self._src >= 0
        np.add.at(outflow, self._src[s], pos[s])
        np.add.at(inflow, self._src[s], neg[s])
        d = self._dst >= 0
        np.add.at(inflow, self._dst[d], pos[d])
        np.add.at(outflow, self._dst[d], neg[d])
        fb_mean = (new["fb"] - old["fb"]) / dt
        f = self._fb_basin >= 0
        np.add.at(inflow, self._fb_basin[f], fb_mean[f])
        rec = dict(
            level=old["level"], storage=old["storage"], inflow_rate=inflow, outflow_rate=outflow,
            precipitation=(new["cp"] - old["cp"]) / dt, surface_runoff=(new["cr"] - old["cr"]) / dt,
            drainage=(new["cd"] - old["cd"]) / dt, evaporation=flow[self._evap],
            infiltration=flow[self._infil], flow=flow, flow_boundary=fb_mean)
        rec["storage_rate"] = (new["storage"] - old["storage"]) / dt
        total_in = inflow + rec["precipitation"] + rec["drainage"] + rec["surface_runoff"]
        total_out = outflow + rec["evaporation"] + rec["infiltration"]
        rec["balance_error"] = rec["storage_rate"] - (total_in - total_out)
        mean_rate = (total_in + total_out) / 2
        with np.errstate(divide="ignore", invalid="ignore"):
            rec["relative_error"] = np.where(mean_rate == 0.0, 0.0, rec["balance_error"] / mean_rate)
        self.periods.append(rec)
        self.times.append(t)
        self._prev = new
        self.subgrid_levels.append(self.m.subgrid.levels(new["level"], t))
        if self.check:
            bad = (np.abs(rec["balance_error"]) > self.abstol) & (np.abs(rec["relative_error"]) > self.reltol)
            if bad.any():
                b, m = np.argwhere(bad)[0]
                raise WaterBalanceError(
                    f"Too large water balance error(s) detected at t = {self._datetime(t)}: "
                    f"Basin #{self.node_id[b]} member {m} balance_error {rec['balance_error'][b, m]:.3e} "
                    f"relative_error {rec['relative_error'][b, m]:.3e}")
        return rec

    # ------------------------------------------------------------------ tables
    def _datetime(self, t: float):
        return self.m.tables.starttime + timedelta(seconds=t)

    def _stack(self, key: str) -> np.ndarray:
        n_ent = self.nb
        return np.stack([r[key] for r in self.periods]) if self.periods else np.zeros((0, n_ent, self.m.B))

    def basin_data(self) -> dict[str, np.ndarray]:
        """`basin_data(model; table=false)`: `[n_periods][n_basin][n_members]` per variable; `time`
        = start of each period (seconds since the start time)."""
        out = dict(time=np.array(self.times[:-1]), node_id=self.node_id)
        for k in BASIN_VARS:
            out[k] = self._stack(k)
        return out

    def flow_data(self) -> dict[str, np.ndarray]:
        """`flow_data(model; table=false)`: mean flow per external link and period, the links over
        a Junction carrying the sum of the internal links routed over them."""
        n_t, B = len(self.periods), self.m.B
        rate = np.zeros((n_t, len(self.link_id), B))
        for k, r in enumerate(self.periods):
            internal = np.zeros((len(self._int_src), B))
            st = self._int_src > 0
            internal[st] = r["flow"][self._int_src[st] - 1]
            internal[~st] = r["flow_boundary"][-self._int_src[~st] - 1]
            rate[k] = self._map @ internal
        return dict(time=np.array(self.times[:-1]), link_id=self.link_id, from_node_id=self.from_node_id,
                    to_node_id=self.to_node_id, flow_rate=rate)

    def subgrid_level_data(self) -> dict[str, np.ndarray]:
        """`subgrid_level_data(model; table=false)` (write.jl:797-817): every save time including
        the start; `[n_times][n_subgrid][n_members]`."""
        return dict(time=np.array(self.times), subgrid_id=self.m.subgrid.subgrid_id,
                    subgrid_level=np.stack(self.subgrid_levels))

    def discrete_control_data(self) -> dict[str, np.ndarray]:
        """`discrete_control_data` (write.jl:565-572) as a table over all members: one row per
        control state change; `realization` tells the member."""
        rows = [(m, *rec) for m in range(self.m.B) for rec in self.m.control_record(m)]
        return dict(realization=np.array([x[0] for x in rows], dtype=np.int32),
                    time=np.array([x[1] for x in rows], dtype=float),
                    control_node_id=np.array([x[2] for x in rows], dtype=np.int32),
                    truth_state=[x[3] for x in rows], control_state=[x[4] for x in rows])

    def _write_control(self, path: Path) -> None:
        from scipy.io import netcdf_file

        d = self.discrete_control_data()
        n = len(d["time"])
        with netcdf_file(str(path), "w") as ds:
            for k, v in CF_GLOBAL.items():
                setattr(ds, k, v)
            ds.title = "Ribasim results: control"
            ds.createDimension("time", n)
            tv = ds.createVariable("time", "d", ("time",))
            tv.units = "seconds since " + self.m.tables.starttime.strftime("%Y-%m-%d %H:%M:%S")
            for k, v in CF["time"].items():
                setattr(tv, k, v)
            tv[:] = d["time"]
            for name in ("control_node_id", "realization"):
                v = ds.createVariable(name, "i", ("time",))
                v[:] = d[name]
            for name in ("truth_state", "control_state"):
                width = max([len(x) for x in d[name]] + [1])
                ds.createDimension(f"{name}_strlen", width)
                v = ds.createVariable(name, "c", ("time", f"{name}_strlen"))
                for k, x in enumerate(d[name]):
                    v[k] = np.frombuffer(x.encode().ljust(width, b"\0"), dtype="S1")

    def basin_state_data(self) -> dict[str, np.ndarray]:
        return dict(node_id=self.node_id, level=self.m.level())

    # ------------------------------------------------------------------ files
    def _write(self, path: Path, title: str, id_name: str, ids: dict[str, np.ndarray],
               time: np.ndarray | None, data: dict[str, np.ndarray]) -> None:
        from scipy.io import netcdf_file

        B = self.m.B
        with netcdf_file(str(path), "w") as ds:
            for k, v in CF_GLOBAL.items():
                setattr(ds, k, v)
            ds.title = f"Ribasim results: {title}"
            n_id = len(ids[id_name])
            ds.createDimension(id_name, n_id)
            dims: tuple[str, ...] = (id_name,)
            if time is not None:
                ds.createDimension("time", len(time))
                tv = ds.createVariable("time", "d", ("time",))
                tv.units = "seconds since " + self.m.tables.starttime.strftime("%Y-%m-%d %H:%M:%S")
                for k, v in CF["time"].items():
                    setattr(tv, k, v)
                tv[:] = time
                dims = ("time", id_name)
            if B > 1:
                ds.createDimension("realization", B)
                rv = ds.createVariable("realization", "i", ("realization",))
                rv.standard_name = "realization"
                rv[:] = np.arange(B, dtype=np.int32)
                dims = ("realization", *dims)
            for name, arr in ids.items():
                v = ds.createVariable(name, "i", (id_name,))
                for k, a in CF[name].items():
                    setattr(v, k, a)
                v[:] = arr
            for name, arr in data.items():
                v = ds.createVariable(name, "d", dims)
                for k, a in CF[name].items():
                    setattr(v, k, a)
                # [time][entity][member] -> ([member])[time][entity]
                a = np.moveaxis(arr, -1, 0) if B > 1 else arr[..., 0]
                if a.size:
                    v[:] = a

    def write(self, results_dir: str | Path) -> dict[str, Path]:
        """`write_results` for the hot path's outputs: basin.nc, flow.nc, basin_state.nc."""
        d = Path(results_dir)
        d.mkdir(parents=True, exist_ok=True)
        paths = {k: d / v for k, v in RESULTS_FILENAME.items()}
        bd = self.basin_data()
        self._write(paths["basin"], "basin", "node_id", dict(node_id=self.node_id), bd["time"],
                    {k: bd[k] for k in BASIN_VARS})
        fd = self.flow_data()
        self._write(paths["flow"], "flow", "link_id",
                    dict(link_id=fd["link_id"], from_node_id=fd["from_node_id"], to_node_id=fd["to_node_id"]),
                    fd["time"], dict(flow_rate=fd["flow_rate"]))
        self._write(paths["basin_state"], "basin_state", "node_id", dict(node_id=self.node_id), None,
                    dict(level=self.basin_state_data()["level"]))
        if len(self.m.subgrid):
            sd = self.subgrid_level_data()
            self._write(paths["subgrid_level"], "subgrid_level", "subgrid_id", dict(subgrid_id=sd["subgrid_id"]),
                        sd["time"], dict(subgrid_level=sd["subgrid_level"]))
        else:
            del paths["subgrid_level"]
        if self.m.flat.ints["n_dc"] > 0:
            self._write_control(paths["control"])
        else:
            del paths["control"]
        return paths
