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

import sqlite3
import tomllib
from dataclasses import dataclass, field
from datetime import datetime
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
    if "." in s:
        s = s.split(".")[0]
    if len(s) == 10:
        s += " 00:00:00"
    return datetime.strptime(s, TIME_FMT)


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
        return m
