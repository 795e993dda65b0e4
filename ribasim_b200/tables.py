"""In-memory model tables and their on-disk form (TOML + GeoPackage/SQLite).

The on-disk contract follows the reference core's reader:
  * `Node(node_id, node_type, subnetwork_id, route_priority, cyclic_time)` and
    `Link(link_id, from_node_id, to_node_id, link_type)` — core/src/graph.jl:10-28,
    core/src/read.jl:1805-1853.
  * one SQLite table per `"<NodeType> / <kind>"` — core/src/config.jl:56-59; columns are
    the struct fields in core/src/schema.jl.
  * rows are sorted on load with the keys in core/src/validation.jl:183-260.
  * TOML keys: core/src/config.jl:180-261.

Only plain `sqlite3` is used (GeoPackage geometry columns are not needed by the core).
"""

from __future__ import annotations

import math
import sqlite3
import tomllib
from dataclasses import dataclass, field
from datetime import datetime, timedelta
from pathlib import Path
from typing import Any

# node types in the order the reference enumerates them (alphabetical; the vertex insertion
# order of the graph is `ORDER BY node_type, node_id`, core/src/graph.jl:12)
NODE_TYPES = [
    "Basin",
    "ContinuousControl",
    "DiscreteControl",
    "FlowBoundary",
    "FlowDemand",
    "Junction",
    "LevelBoundary",
    "LevelDemand",
    "LinearResistance",
    "ManningResistance",
    "Observation",
    "Outlet",
    "PidControl",
    "Pump",
    "TabulatedRatingCurve",
    "Terminal",
    "UserDemand",
]

CONTROL_SOURCE_TYPES = {
    "DiscreteControl",
    "ContinuousControl",
    "PidControl",
    "FlowDemand",
    "LevelDemand",
}

# table -> (columns, sort key columns).  Column types: f=float, i=int, s=str, t=time, b=bool
SCHEMA: dict[str, tuple[dict[str, str], tuple[str, ...]]] = {
    "Basin / static": (
        dict(node_id="i", drainage="f", potential_evaporation="f", infiltration="f",
             precipitation="f", surface_runoff="f"),
        ("node_id",),
    ),
    "Basin / time": (
        dict(node_id="i", time="t", drainage="f", potential_evaporation="f", infiltration="f",
             precipitation="f", surface_runoff="f"),
        ("node_id", "time"),
    ),
    "Basin / profile": (
        dict(node_id="i", area="f", level="f", storage="f"),
        ("node_id", "level"),
    ),
    "Basin / state": (dict(node_id="i", level="f"), ("node_id",)),
    "Basin / subgrid": (
        dict(subgrid_id="i", node_id="i", basin_level="f", subgrid_level="f"),
        ("subgrid_id", "basin_level"),
    ),
    "Basin / subgrid_time": (
        dict(subgrid_id="i", node_id="i", time="t", basin_level="f", subgrid_level="f"),
        ("subgrid_id", "time", "basin_level"),
    ),
    "Pump / static": (
        dict(node_id="i", flow_rate="f", min_flow_rate="f", max_flow_rate="f",
             min_upstream_level="f", max_downstream_level="f", control_state="s",
             allocation_controlled="b"),
        ("node_id", "control_state"),
    ),
    "Pump / time": (
        dict(node_id="i", time="t", flow_rate="f", min_flow_rate="f", max_flow_rate="f",
             min_upstream_level="f", max_downstream_level="f"),
        ("node_id", "time"),
    ),
    "Outlet / static": (
        dict(node_id="i", flow_rate="f", min_flow_rate="f", max_flow_rate="f",
             min_upstream_level="f", max_downstream_level="f", control_state="s",
             allocation_controlled="b"),
        ("node_id", "control_state"),
    ),
    "Outlet / time": (
        dict(node_id="i", time="t", flow_rate="f", min_flow_rate="f", max_flow_rate="f",
             min_upstream_level="f", max_downstream_level="f"),
        ("node_id", "time"),
    ),
    "LevelBoundary / static": (dict(node_id="i", level="f"), ("node_id",)),
    "LevelBoundary / time": (dict(node_id="i", time="t", level="f"), ("node_id", "time")),
    "FlowBoundary / static": (dict(node_id="i", flow_rate="f"), ("node_id",)),
    "FlowBoundary / time": (dict(node_id="i", time="t", flow_rate="f"), ("node_id", "time")),
    "LinearResistance / static": (
        dict(node_id="i", resistance="f", max_flow_rate="f", control_state="s"),
        ("node_id", "control_state"),
    ),
    "ManningResistance / static": (
        dict(node_id="i", length="f", manning_n="f", profile_width="f", profile_slope="f",
             control_state="s"),
        ("node_id", "control_state"),
    ),
    "TabulatedRatingCurve / static": (
        dict(node_id="i", level="f", flow_rate="f", max_downstream_level="f",
             control_state="s", allocation_controlled="b"),
        ("node_id", "control_state", "level"),
    ),
    "TabulatedRatingCurve / time": (
        dict(node_id="i", time="t", level="f", flow_rate="f", max_downstream_level="f"),
        ("node_id", "time", "level"),
    ),
    "DiscreteControl / variable": (
        dict(node_id="i", compound_variable_id="i", listen_node_id="i", variable="s",
             weight="f", look_ahead="f"),
        ("node_id", "compound_variable_id", "listen_node_id", "variable"),
    ),
    "DiscreteControl / condition": (
        dict(node_id="i", compound_variable_id="i", condition_id="i", threshold_high="f",
             threshold_low="f", time="t"),
        ("node_id", "compound_variable_id", "condition_id"),
    ),
    "DiscreteControl / logic": (
        dict(node_id="i", truth_state="s", control_state="s"),
        ("node_id", "truth_state"),
    ),
    "ContinuousControl / variable": (
        dict(node_id="i", listen_node_id="i", variable="s", weight="f", look_ahead="f"),
        ("node_id", "listen_node_id", "variable"),
    ),
    "ContinuousControl / function": (
        dict(node_id="i", input="f", output="f", controlled_variable="s"),
        ("node_id", "input"),
    ),
    "PidControl / static": (
        dict(node_id="i", listen_node_id="i", target="f", proportional="f", integral="f",
             derivative="f", control_state="s"),
        ("node_id", "control_state"),
    ),
    "PidControl / time": (
        dict(node_id="i", listen_node_id="i", time="t", target="f", proportional="f",
             integral="f", derivative="f"),
        ("node_id", "time"),
    ),
    "UserDemand / static": (
        dict(node_id="i", demand="f", return_factor="f", min_level="f", demand_priority="i"),
        ("node_id", "demand_priority"),
    ),
    "UserDemand / time": (
        dict(node_id="i", time="t", demand="f", return_factor="f", min_level="f",
             demand_priority="i"),
        ("node_id", "demand_priority", "time"),
    ),
    "FlowDemand / static": (
        dict(node_id="i", demand="f", demand_priority="i"),
        ("node_id", "demand_priority"),
    ),
    "FlowDemand / time": (
        dict(node_id="i", time="t", demand="f", demand_priority="i"),
        ("node_id", "demand_priority", "time"),
    ),
}

TIME_FMT = "%Y-%m-%d %H:%M:%S"

SOLVER_DEFAULTS: dict[str, Any] = dict(
    algorithm="QNDF",
    saveat=86400.0,
    dt=None,
    dtmin=0.0,
    dtmax=None,
    force_dtmin=False,
    abstol=1.0e-5,
    reltol=1.0e-5,
    water_balance_abstol=1.0e-3,
    water_balance_reltol=1.0e-2,
    maxiters=int(1e9),
    sparse=True,
    autodiff=True,
    evaporate_mass=True,
    depth_threshold=0.1,
    level_difference_threshold=0.02,
    specialize=False,
)

INTERPOLATION_DEFAULTS: dict[str, Any] = dict(flow_boundary="block", block_transition_period=0.0)


def parse_time(x: Any) -> datetime:
    """Accept datetime or the string forms the core accepts (core/src/read.jl:2099-2108)."""
    if isinstance(x, datetime):
        return x
    s = str(x).replace("T", " ")
    ms = 0
    if "." in s:
        s, frac = s.split(".", 1)
        # extra fractional digits are truncated to milliseconds (parse_time_strings!, read.jl)
        ms = int((frac + "000")[:3])
    if len(s) == 10:
        s += " 00:00:00"
    return datetime.strptime(s, TIME_FMT) + timedelta(milliseconds=ms)


def datetime_since(t: float, t0: datetime) -> datetime:
    """`datetime_since` (core/src/util.jl): seconds since t0 -> datetime, rounded to milliseconds."""
    return t0 + timedelta(milliseconds=round(1000.0 * t))


def seconds_since(t: datetime, t0: datetime) -> float:
    """`seconds_since` (core/src/util.jl): milliseconds resolution."""
    return 0.001 * round((t - t0) / timedelta(milliseconds=1))


def snake_case(name: str) -> str:
    """CamelCase -> snake_case (core/src/config.jl:134-138)."""
    import re
    return re.sub(r"(?<!^)(?=[A-Z])", "_", name).lower()


NC_DIM_NAMES = ("time", "node_id", "link_id", "subgrid_id", "substance", "demand_priority")
_CF_UNIT_SECONDS = dict(seconds=1.0, second=1.0, s=1.0, minutes=60.0, minute=60.0, hours=3600.0,
                        hour=3600.0, days=86400.0, day=86400.0, milliseconds=1e-3)


def _nc_attr(var, name: str) -> str | None:
    v = getattr(var, name, None)
    if v is None:
        return None
    return v.decode() if isinstance(v, bytes) else str(v)


def load_netcdf(path: str | Path, table_name: str) -> list[dict[str, Any]]:
    """A table stored as dense arrays in a NetCDF file -> rows (`load_netcdf`,
    core/src/read.jl:1889-1984; coordinate detection `:1986-2050`).

    Every table has a node_id dimension (`cf_role = "timeseries_id"` or name `node_id`, integers
    or the NetCDF3 char matrix Delft-FEWS writes), some a time (`axis = "T"`, `standard_name =
    "time"` or name `time`) and a priority dimension (`realization`).  Data variables are stored
    with dimensions (time, [priority,] node) in file order; rows whose data variables are all
    missing are dropped.  Read with scipy's classic-format reader (NetCDF4/HDF5 files need
    netCDF4, which is not a dependency)."""
    import numpy as np
    from scipy.io import netcdf_file

    cols, _ = SCHEMA[table_name]
    data_names = [c for c in cols if c not in NC_DIM_NAMES]
    with netcdf_file(str(path), "r", mmap=False) as ds:
        def find(pred):
            for name, var in ds.variables.items():
                if pred(name, var):
                    return name, var
            return None, None

        _, id_var = find(lambda n, v: _nc_attr(v, "cf_role") == "timeseries_id" or n == "node_id")
        if id_var is None:
            raise ValueError(f"{path}: no node_id coordinate (cf_role = timeseries_id)")
        ids = np.array(id_var[:])
        if ids.dtype.kind == "S" and ids.ndim == 2:  # (n_id, n_char) char matrix, NUL padded
            node_id = np.array([int(b"".join(row).split(b"\0")[0].decode().strip()) for row in ids],
                               dtype=np.int64)
        else:
            node_id = ids.astype(np.int64).reshape(-1)
        _, t_var = find(lambda n, v: _nc_attr(v, "axis") == "T"
                        or _nc_attr(v, "standard_name") == "time" or n == "time")
        _, p_var = find(lambda n, v: _nc_attr(v, "standard_name") == "realization"
                        or n == "realization")
        n_id = len(node_id)
        times = None
        if t_var is not None:
            units = _nc_attr(t_var, "units") or ""
            unit, _, since = units.partition(" since ")
            if unit.strip() not in _CF_UNIT_SECONDS or not since:
                raise ValueError(f"{path}: unsupported time units `{units}`")
            t0 = parse_time(since.strip().replace("Z", "").split("+")[0].strip())
            from datetime import timedelta
            times = [t0 + timedelta(seconds=float(x) * _CF_UNIT_SECONDS[unit.strip()])
                     for x in np.array(t_var[:]).reshape(-1)]
        prio = None if p_var is None else np.array(p_var[:]).astype(np.int64).reshape(-1)
        n_t = 1 if times is None else len(times)
        n_p = 1 if prio is None else len(prio)
        n = n_id * n_p * n_t
        columns: dict[str, Any] = {}
        for name in data_names:
            if name not in ds.variables:
                continue
            var = ds.variables[name]
            arr = np.array(var[:], dtype=np.float64)
            for att in ("_FillValue", "missing_value"):
                fv = getattr(var, att, None)
                if fv is not None:
                    arr = np.where(arr == np.float64(np.array(fv).reshape(-1)[0]), np.nan, arr)
            if table_name == "UserDemand / time":
                # demand (time, priority, node), return_factor (time, node), min_level (node)
                if arr.ndim == 1:
                    arr = np.broadcast_to(arr[None, None, :], (n_t, n_p, n_id))
                elif arr.ndim == 2:
                    arr = np.broadcast_to(arr[:, None, :], (n_t, n_p, n_id))
            flat = np.ascontiguousarray(arr).reshape(-1)
            if flat.size != n:
                raise ValueError(f"Inconsistent lengths in NetCDF file {path} for variable {name}")
            columns[name] = flat
    rows: list[dict[str, Any]] = []
    for k in range(n):  # node fastest, then priority, then time
        i, rest = k % n_id, k // n_id
        pk, tk = rest % n_p, rest // n_p
        vals = {c: float(v[k]) for c, v in columns.items()}
        if vals and all(math.isnan(x) for x in vals.values()):
            continue
        row: dict[str, Any] = {c: None for c in cols}
        row["node_id"] = int(node_id[i])
        if times is not None and "time" in cols:
            row["time"] = times[tk]
        if prio is not None and "demand_priority" in cols:
            row["demand_priority"] = int(prio[pk])
        for c, x in vals.items():
            if math.isnan(x):
                row[c] = None
            elif cols[c] == "i":
                row[c] = int(x)
            elif cols[c] == "b":
                row[c] = bool(x)
            else:
                row[c] = x
        rows.append(row)
    return rows


def write_netcdf(path: str | Path, table_name: str, rows: list[dict[str, Any]],
                 starttime: datetime, char_ids: bool = False) -> None:
    """Write a time table as the dense (time, node) NetCDF the core reads (classic format);
    `char_ids` stores the ids as the Delft-FEWS char matrix."""
    import numpy as np
    from scipy.io import netcdf_file

    cols, _ = SCHEMA[table_name]
    data_names = [c for c in cols if c not in NC_DIM_NAMES and cols[c] in ("f", "i")]
    ids = sorted({int(r["node_id"]) for r in rows})
    times = sorted({parse_time(r["time"]) for r in rows}) if "time" in cols else []
    with netcdf_file(str(path), "w") as ds:
        ds.createDimension("stations", len(ids))
        if char_ids:
            ds.createDimension("char_leng_id", 16)
            v = ds.createVariable("station_id", "c", ("stations", "char_leng_id"))
            v.cf_role = "timeseries_id"
            for i, x in enumerate(ids):
                txt = str(x).encode().ljust(16, b"\0")
                v[i] = np.frombuffer(txt, dtype="S1")
        else:
            v = ds.createVariable("node_id", "i", ("stations",))
            v.cf_role = "timeseries_id"
            v[:] = np.array(ids, dtype=np.int32)
        dims = ("stations",)
        if times:
            ds.createDimension("time", len(times))
            tv = ds.createVariable("time", "d", ("time",))
            tv.units = "seconds since " + starttime.strftime(TIME_FMT)
            tv.axis = "T"
            tv[:] = np.array([(t - starttime).total_seconds() for t in times])
            dims = ("time", "stations")
        ii = {x: k for k, x in enumerate(ids)}
        ti = {t: k for k, t in enumerate(times)}
        for name in data_names:
            arr = np.full((len(times) or 1, len(ids)), np.nan)
            for r in rows:
                if r.get(name) is not None:
                    arr[ti[parse_time(r["time"])] if times else 0, ii[int(r["node_id"])]] = float(r[name])
            var = ds.createVariable(name, "d", dims)
            var._FillValue = np.float64(-999.0)
            arr = np.where(np.isnan(arr), -999.0, arr)
            var[:] = arr if times else arr[0]


def _sort_key(cols: tuple[str, ...]):
    # `missing` sorts after every value in the reference (Julia `isless`).
    def key(row: dict[str, Any]):
        out = []
        for c in cols:
            v = row.get(c)
            if c == "time" and v is not None:
                v = parse_time(v)
            out.append((v is None, v if v is not None else 0))
        return tuple(out)

    return key


@dataclass
class ModelTables:
    """Plain container of everything the core reads from `ribasim.toml` + `database.gpkg`."""

    starttime: datetime
    endtime: datetime
    crs: str = "EPSG:28992"
    ribasim_version: str = "2026.1.2"
    input_dir: str = "input"
    results_dir: str = "results"
    solver: dict[str, Any] = field(default_factory=dict)
    interpolation: dict[str, Any] = field(default_factory=dict)
    experimental: dict[str, Any] = field(default_factory=dict)
    node: list[dict[str, Any]] = field(default_factory=list)
    link: list[dict[str, Any]] = field(default_factory=list)
    tables: dict[str, list[dict[str, Any]]] = field(default_factory=dict)
    external_tables: dict[str, str] = field(default_factory=dict)  # table -> NetCDF path in input_dir

    # ------------------------------------------------------------------ builder
    def add_node(self, node_type: str, node_id: int, *, subnetwork_id: int | None = None,
                 route_priority: int | None = None, cyclic_time: bool = False,
                 **tables: Any) -> int:
        """Add a node.  Each keyword is a table kind (`static`, `profile`, ...) whose value is
        a dict of columns (scalars are broadcast against list-valued columns)."""
        if node_type not in NODE_TYPES:
            raise ValueError(f"unknown node type {node_type}")
        self.node.append(dict(node_id=int(node_id), node_type=node_type,
                              subnetwork_id=subnetwork_id, route_priority=route_priority,
                              cyclic_time=bool(cyclic_time)))
        for kind, cols in tables.items():
            name = f"{node_type} / {kind}"
            if name not in SCHEMA:
                raise ValueError(f"unknown table {name}")
            n = 1
            for v in cols.values():
                if isinstance(v, (list, tuple)) or hasattr(v, "__len__") and not isinstance(v, str):
                    n = max(n, len(v))
            rows = self.tables.setdefault(name, [])
            for k in range(n):
                row: dict[str, Any] = dict(node_id=int(node_id))
                for c, v in cols.items():
                    if c not in SCHEMA[name][0]:
                        raise ValueError(f"unknown column {c} in {name}")
                    if hasattr(v, "__len__") and not isinstance(v, str):
                        val = v[k]
                    else:
                        val = v
                    if hasattr(val, "item"):
                        val = val.item()
                    row[c] = val
                rows.append(row)
        return int(node_id)

    def node_type_of(self, node_id: int) -> str:
        for r in self.node:
            if r["node_id"] == node_id:
                return r["node_type"]
        raise KeyError(node_id)

    def add_link(self, from_node_id: int, to_node_id: int, link_type: str | None = None,
                 link_id: int | None = None) -> int:
        if link_type is None:
            link_type = (
                "control" if self.node_type_of(from_node_id) in CONTROL_SOURCE_TYPES else "flow"
            )
        if link_id is None:
            link_id = 1 + max((r["link_id"] for r in self.link), default=0)
        self.link.append(dict(link_id=int(link_id), from_node_id=int(from_node_id),
                              to_node_id=int(to_node_id), link_type=link_type))
        return link_id

    # ------------------------------------------------------------------ access
    def table(self, name: str) -> list[dict[str, Any]]:
        """Rows of a table in the reference's required sort order (validation.jl:183-260)."""
        rows = self.tables.get(name, [])
        cols, keys = SCHEMA[name]
        full = [{c: r.get(c) for c in cols} for r in rows]
        return sorted(full, key=_sort_key(keys))

    def solver_option(self, key: str) -> Any:
        return self.solver.get(key, SOLVER_DEFAULTS[key])

    def seconds_since(self, t: Any) -> float:
        return (parse_time(t) - self.starttime).total_seconds()

    @property
    def t_end(self) -> float:
        return (self.endtime - self.starttime).total_seconds()

    # ------------------------------------------------------------------ disk
    def write(self, directory: str | Path) -> Path:
        """Write `ribasim.toml` + `<input_dir>/database.gpkg`; returns the TOML path."""
        directory = Path(directory)
        (directory / self.input_dir).mkdir(parents=True, exist_ok=True)
        toml_path = directory / "ribasim.toml"
        lines = [
            f"starttime = {self.starttime.isoformat()}",
            f"endtime = {self.endtime.isoformat()}",
            f'crs = "{self.crs}"',
            f'ribasim_version = "{self.ribasim_version}"',
            f'input_dir = "{self.input_dir}"',
            f'results_dir = "{self.results_dir}"',
        ]
        for section, d in (("solver", self.solver), ("interpolation", self.interpolation),
                           ("experimental", self.experimental)):
            if d:
                lines.append("")
                lines.append(f"[{section}]")
                for k, v in d.items():
                    if v is None:
                        continue
                    if isinstance(v, bool):
                        lines.append(f"{k} = {'true' if v else 'false'}")
                    elif isinstance(v, str):
                        lines.append(f'{k} = "{v}"')
                    else:
                        lines.append(f"{k} = {v!r}")
        by_section: dict[str, list[str]] = {}
        for name, rel in self.external_tables.items():
            node_type, kind = name.split(" / ")
            by_section.setdefault(snake_case(node_type), []).append(f'{kind} = "{rel}"')
            write_netcdf(directory / self.input_dir / rel, name, self.table(name), self.starttime)
        for section, entries in by_section.items():
            lines.append("")
            lines.append(f"[{section}]")
            lines.extend(entries)
        toml_path.write_text("\n".join(lines) + "\n")

        db_path = directory / self.input_dir / "database.gpkg"
        if db_path.exists():
            db_path.unlink()
        con = sqlite3.connect(db_path)
        try:
            con.execute(
                "CREATE TABLE Node (node_id INTEGER PRIMARY KEY, node_type TEXT NOT NULL, "
                "subnetwork_id INTEGER, route_priority INTEGER, cyclic_time INTEGER)"
            )
            con.executemany(
                "INSERT INTO Node VALUES (?,?,?,?,?)",
                [(r["node_id"], r["node_type"], r["subnetwork_id"], r["route_priority"],
                  int(r["cyclic_time"])) for r in self.node],
            )
            con.execute(
                "CREATE TABLE Link (link_id INTEGER PRIMARY KEY, from_node_id INTEGER, "
                "to_node_id INTEGER, link_type TEXT)"
            )
            con.executemany(
                "INSERT INTO Link VALUES (?,?,?,?)",
                [(r["link_id"], r["from_node_id"], r["to_node_id"], r["link_type"])
                 for r in self.link],
            )
            sqltype = dict(f="REAL", i="INTEGER", s="TEXT", t="TEXT", b="INTEGER")
            for name in self.tables:
                if name in self.external_tables:
                    continue
                cols, _ = SCHEMA[name]
                decl = ", ".join(f'"{c}" {sqltype[k]}' for c, k in cols.items())
                con.execute(f'CREATE TABLE "{name}" ({decl})')
                rows = []
                for r in self.table(name):
                    vals = []
                    for c, k in cols.items():
                        v = r[c]
                        if v is not None and k == "t":
                            v = parse_time(v).strftime(TIME_FMT)
                        if v is not None and k == "b":
                            v = int(v)
                        vals.append(v)
                    rows.append(tuple(vals))
                con.executemany(
                    f'INSERT INTO "{name}" VALUES ({",".join("?" * len(cols))})', rows
                )
            con.commit()
        finally:
            con.close()
        return toml_path

    @classmethod
    def read(cls, toml_path: str | Path) -> "ModelTables":
        """Load a model from `ribasim.toml` (+ the GeoPackage it points at)."""
        toml_path = Path(toml_path)
        with open(toml_path, "rb") as f:
            cfg = tomllib.load(f)
        for key in ("starttime", "endtime", "crs", "ribasim_version", "input_dir", "results_dir"):
            if key not in cfg:
                raise ValueError(f"ribasim.toml: missing required key `{key}`")
        m = cls(
            starttime=parse_time(cfg["starttime"]),
            endtime=parse_time(cfg["endtime"]),
            crs=cfg["crs"],
            ribasim_version=cfg["ribasim_version"],
            input_dir=cfg["input_dir"],
            results_dir=cfg["results_dir"],
            solver=dict(cfg.get("solver", {})),
            interpolation=dict(cfg.get("interpolation", {})),
            experimental=dict(cfg.get("experimental", {})),
        )
        db_path = toml_path.parent / m.input_dir / "database.gpkg"
        if not db_path.is_file():
            raise FileNotFoundError(f"Database file not found: {db_path}")
        con = sqlite3.connect(db_path)
        try:
            con.row_factory = sqlite3.Row
            have = {r[0] for r in con.execute("SELECT name FROM sqlite_master WHERE type='table'")}
            node_cols = {r[1] for r in con.execute("PRAGMA table_info(Node)")}
            for r in con.execute("SELECT * FROM Node"):
                m.node.append(dict(
                    node_id=int(r["node_id"]), node_type=r["node_type"],
                    subnetwork_id=r["subnetwork_id"] if "subnetwork_id" in node_cols else None,
                    route_priority=r["route_priority"] if "route_priority" in node_cols else None,
                    cyclic_time=bool(r["cyclic_time"]) if "cyclic_time" in node_cols
                    and r["cyclic_time"] is not None else False,
                ))
            for r in con.execute("SELECT * FROM Link"):
                m.link.append(dict(link_id=int(r["link_id"]), from_node_id=int(r["from_node_id"]),
                                   to_node_id=int(r["to_node_id"]), link_type=r["link_type"]))
            for name, (cols, _) in SCHEMA.items():
                if name not in have:
                    continue
                present = {r[1] for r in con.execute(f'PRAGMA table_info("{name}")')}
                rows = []
                for r in con.execute(f'SELECT * FROM "{name}"'):
                    row = {}
                    for c, k in cols.items():
                        v = r[c] if c in present else None
                        if v is not None and k == "b":
                            v = bool(v)
                        row[c] = v
                    rows.append(row)
                if rows:
                    m.tables[name] = rows
        finally:
            con.close()
        # tables stored outside the database: `[basin] time = "basin_time.nc"` (load_data,
        # core/src/read.jl:2057-2098; only NetCDF is supported there too)
        for name in SCHEMA:
            node_type, kind = name.split(" / ")
            path = cfg.get(snake_case(node_type), {}).get(kind) if isinstance(cfg.get(snake_case(node_type)), dict) else None
            if path is None:
                continue
            table_path = toml_path.parent / m.input_dir / path
            if table_path.suffix.lower() != ".nc":
                raise ValueError(f"Unsupported file format: {table_path}. Only '.nc' is supported.")
            m.tables[name] = load_netcdf(table_path, name)
            m.external_tables[name] = str(path)
        return m
