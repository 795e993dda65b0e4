"""Host network compiler: model tables -> parameters -> flat arrays for the CUDA library.

This is the host-side mirror of `Ribasim.Parameters(db, config)` (core/src/read.jl:1692-1802)
restricted to what the implicit-integration hot path needs.  It reproduces the reference's
integer layout rules bit-exactly:

* graph vertex codes: nodes inserted `ORDER BY node_type, node_id`, Observation skipped
  (core/src/graph.jl:10-13, 56-75); Junction elimination with `rem_vertex!`
  swap-with-last renumbering (core/src/graph.jl:140-211);
* node index = rank of node_id inside its type (core/src/read.jl:1819-1848);
* state vector order `state_components` (core/src/parameter.jl:11-22,
  core/src/util.jl:694-717);
* basin inflow/outflow neighbour order = ascending vertex code (Graphs.jl adjacency lists),
  which fixes the summation order of `reduce_state!` (core/src/util.jl:745-791).
"""

from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Any

import numpy as np

from .interp import (
    IndexLookup,
    LinearTable,
    PchipTable,
    TimeSeries,
    as_f64,
    as_i32,
    finite_difference,
    storage_to_level,
    trapz_integrate,
)
from .tables import INTERPOLATION_DEFAULTS, ModelTables

# endpoint kinds understood by `get_level` (core/src/util.jl:170-191)
KIND_BASIN, KIND_LEVEL_BOUNDARY, KIND_TERMINAL, KIND_NONE = 0, 1, 2, 3
# connector state types in state order (core/src/parameter.jl:11-22)
CT_TRC, CT_PUMP, CT_OUTLET, CT_UD_IN, CT_UD_OUT, CT_LR, CT_MANNING = range(7)
# continuous control types (core/src/parameter.jl:37)
CONTROL_NONE, CONTROL_CONTINUOUS, CONTROL_PID = 0, 1, 2
# compound-variable subvariable kinds (core/src/util.jl:491-531, callback.jl:718-755)
SV_BASIN_LEVEL, SV_BASIN_STORAGE, SV_DU_STATE, SV_CONSTANT = 0, 1, 2, 3

STATE_COMPONENTS = (
    "tabulated_rating_curve",
    "pump",
    "outlet",
    "user_demand_inflow",
    "user_demand_outflow",
    "linear_resistance",
    "manning_resistance",
    "evaporation",
    "infiltration",
    "integral",
)

CONSERVATIVE = {"Pump", "Outlet", "TabulatedRatingCurve", "LinearResistance", "ManningResistance"}


@dataclass(frozen=True)
class NodeID:
    type: str
    value: int
    idx: int  # 1-based index inside its node type, as in the reference

    def __repr__(self) -> str:
        return f"{self.type} #{self.value}"


@dataclass
class LinkMeta:
    id: int
    type: str
    link: tuple[NodeID, NodeID]


class Graph:
    """Directed meta graph with Graphs.jl/MetaGraphsNext vertex-code semantics."""

    def __init__(self) -> None:
        self.code_of: dict[NodeID, int] = {}
        self.label_of: dict[int, NodeID] = {}
        self.out: dict[int, set[int]] = {}
        self.inn: dict[int, set[int]] = {}
        self.edge: dict[tuple[NodeID, NodeID], LinkMeta] = {}
        self.flow_link_pairs: list[tuple[int, int]] = []  # (external id, internal id), graph.jl:205-208

    def add_vertex(self, label: NodeID) -> None:
        code = len(self.code_of) + 1
        self.code_of[label] = code
        self.label_of[code] = label
        self.out[code] = set()
        self.inn[code] = set()

    def has_edge(self, a: NodeID, b: NodeID) -> bool:
        return (a, b) in self.edge

    def add_edge(self, a: NodeID, b: NodeID, meta: LinkMeta) -> None:
        self.edge[(a, b)] = meta
        ca, cb = self.code_of[a], self.code_of[b]
        self.out[ca].add(cb)
        self.inn[cb].add(ca)

    def rem_vertex(self, label: NodeID) -> None:
        """MetaGraphsNext.rem_vertex!: the last vertex takes over the removed code."""
        v = self.code_of[label]
        n = len(self.code_of)
        for key in [k for k in self.edge if label in k]:
            del self.edge[key]
        for s in list(self.inn[v]):
            self.out[s].discard(v)
        for d in list(self.out[v]):
            self.inn[d].discard(v)
        self.inn[v], self.out[v] = set(), set()
        if v != n:
            last = self.label_of[n]
            ins = self.inn[n] - {n}
            outs = self.out[n] - {n}
            for s in ins:
                self.out[s].discard(n)
                self.out[s].add(v)
            for d in outs:
                self.inn[d].discard(n)
                self.inn[d].add(v)
            self.inn[v], self.out[v] = set(ins), set(outs)
            self.label_of[v] = last
            self.code_of[last] = v
        del self.code_of[label]
        del self.label_of[n]
        del self.inn[n]
        del self.out[n]

    def inneighbors(self, label: NodeID, link_type: str | None = None) -> list[NodeID]:
        out = []
        for c in sorted(self.inn[self.code_of[label]]):
            src = self.label_of[c]
            if link_type is None or self.edge[(src, label)].type == link_type:
                out.append(src)
        return out

    def outneighbors(self, label: NodeID, link_type: str | None = None) -> list[NodeID]:
        out = []
        for c in sorted(self.out[self.code_of[label]]):
            dst = self.label_of[c]
            if link_type is None or self.edge[(label, dst)].type == link_type:
                out.append(dst)
        return out

    def inflow_ids(self, label: NodeID) -> list[NodeID]:
        return self.inneighbors(label, "flow")

    def outflow_ids(self, label: NodeID) -> list[NodeID]:
        return self.outneighbors(label, "flow")


# --------------------------------------------------------------------------- parameters
@dataclass
class BasinP:
    node_id: list[NodeID]
    inflow_ids: list[list[NodeID]]
    outflow_ids: list[list[NodeID]]
    storage0: np.ndarray
    low_storage_threshold: np.ndarray
    # profile tables, one entry per basin
    levels: list[list[float]]
    areas: list[list[float]]  # level_to_area.u
    dSdh: list[list[float]]  # storage_to_level.itp.u
    dSdh_slope: list[list[float]]
    cumS: list[list[float]]  # storage_to_level.t
    area_slope: list[list[float]]
    forcing: dict[str, list[TimeSeries]]
    # current vertical fluxes (BMI-visible)
    vertical_flux: dict[str, np.ndarray]


@dataclass
class Params:
    """Python mirror of `ParametersIndependent` for the node types on the hot path."""

    tables: ModelTables
    graph: Graph
    node_ids_all: list[NodeID]
    basin: BasinP
    nodes: dict[str, dict[str, Any]]
    state_node_id: list[NodeID]
    state_ranges: dict[str, tuple[int, int]]  # 0-based [start, stop)
    state_inflow_link: list[LinkMeta | None]
    state_outflow_link: list[LinkMeta | None]
    link_to_state_idx: dict[tuple[NodeID, NodeID], int]
    internal_flow_links: list[LinkMeta]
    external_flow_links: list[LinkMeta]
    level_difference_threshold: float
    tstops: list[float] = field(default_factory=list)

    @property
    def n_states(self) -> int:
        return len(self.state_node_id)


def _jl(x: float) -> str:
    """A float as the reference prints it in its messages (shortest digits, exponent form from
    1e6 and below 1e-4)."""
    x = float(x)
    if x != 0.0 and math.isfinite(x) and (abs(x) >= 1e6 or abs(x) < 1e-4):
        mant, exp = np.format_float_scientific(x, trim="0").split("e")
        return f"{mant}e{int(exp)}"
    return repr(x)


def _endpoint(nid: NodeID | None) -> tuple[int, int]:
    if nid is None:
        return KIND_NONE, 0
    if nid.type == "Basin":
        return KIND_BASIN, nid.idx - 1
    if nid.type == "LevelBoundary":
        return KIND_LEVEL_BOUNDARY, nid.idx - 1
    if nid.type == "Terminal":
        return KIND_TERMINAL, nid.idx - 1
    return KIND_NONE, nid.idx - 1


def create_graph(m: ModelTables) -> tuple[Graph, list[NodeID], list[LinkMeta], list[LinkMeta]]:
    """core/src/graph.jl:8-211."""
    # get_node_ids(db): ORDER BY node_id with a per-type counter (read.jl:1819-1828)
    rows = sorted(m.node, key=lambda r: r["node_id"])
    ids = [r["node_id"] for r in rows]
    if len(set(ids)) != len(ids):
        raise ValueError("Node IDs must be unique")
    counter: dict[str, int] = {}
    node_ids_all: list[NodeID] = []
    by_value: dict[int, NodeID] = {}
    for r in rows:
        counter[r["node_type"]] = counter.get(r["node_type"], 0) + 1
        nid = NodeID(r["node_type"], r["node_id"], counter[r["node_type"]])
        node_ids_all.append(nid)
        by_value[nid.value] = nid

    g = Graph()
    for r in sorted(m.node, key=lambda r: (r["node_type"], r["node_id"])):
        nid = by_value[r["node_id"]]
        if nid.type == "Observation":
            continue
        g.add_vertex(nid)

    external: list[LinkMeta] = []
    max_link_id = 0
    for r in m.link:
        lt = r["link_type"]
        if lt not in ("flow", "control", "listen", "observation", "none"):
            raise ValueError(f"Invalid link type '{lt}' for link #{r['link_id']} from node "
                             f"#{r['from_node_id']} to node #{r['to_node_id']}.")
        if r["from_node_id"] not in by_value or r["to_node_id"] not in by_value:
            raise ValueError(f"Link {r['link_id']} refers to a node that is not in the Node table")
        src, dst = by_value[r["from_node_id"]], by_value[r["to_node_id"]]
        meta = LinkMeta(r["link_id"], lt, (src, dst))
        if lt == "flow":
            external.append(meta)
        max_link_id = max(max_link_id, r["link_id"])
        if lt in ("listen", "observation"):
            continue
        if g.has_edge(src, dst):
            raise ValueError(f"Duplicate link {src} -> {dst}")
        if g.has_edge(dst, src) and "UserDemand" not in (src.type, dst.type):
            raise ValueError(
                f"Invalid link: the opposite link already exists (only allowed for UserDemand): "
                f"{src} -> {dst}"
            )
        g.add_edge(src, dst, meta)

    # simplify_graph!: remove junctions (graph.jl:140-211)
    internal = [e for e in external if e.link[0].type != "Junction" and e.link[1].type != "Junction"]
    # flow_link_map: (external link id, internal link id) pairs; an external link carries the sum
    # of the internal links routed over it (graph.jl:155-208)
    pairs = [(e.id, e.id) for e in internal]
    link_mapping: dict[int, list[int]] = {}
    new_id = max_link_id + 1
    junctions = sorted((n for n in node_ids_all if n.type == "Junction"), key=lambda n: n.value)
    for j in junctions:
        for a in g.inneighbors(j):
            for b in g.outneighbors(j):
                lid = g.edge[(a, j)].id
                ext_ids = list(link_mapping.get(lid, [lid]))
                lid = g.edge[(j, b)].id
                ext_ids += link_mapping.get(lid, [lid])
                meta = LinkMeta(new_id, "flow", (a, b))
                if a.type != "Junction" and b.type != "Junction":
                    pairs += [(x, new_id) for x in ext_ids]
                    internal.append(meta)
                link_mapping[new_id] = ext_ids
                new_id += 1
                if g.has_edge(a, b):
                    raise ValueError(f"Duplicate link: Junction links form cycle: {a} -> {b}")
                g.add_edge(a, b, meta)
        g.rem_vertex(j)
    g.flow_link_pairs = pairs
    return g, node_ids_all, internal, external


def _group(rows: list[dict], key: str) -> dict[Any, list[dict]]:
    out: dict[Any, list[dict]] = {}
    for r in rows:
        out.setdefault(r[key], []).append(r)
    return out


def _by_node(m: ModelTables, name: str) -> dict[int, list[dict]]:
    """Sorted rows of table `name` grouped by node_id (memoised on the model)."""
    cache = m.__dict__.setdefault("_by_node_cache", {})
    if name not in cache:
        cache[name] = _group(m.table(name), "node_id") if name in m.tables else {}
    return cache[name]


def _coalesce(v, default):
    return default if v is None else v


def _parse_series(m: ModelTables, node_type: str, node_ids: list[NodeID], param: str, *,
                  default=None, is_complete=True, kind="constant",
                  cyclic: dict[int, bool] | None = None, time_table=True,
                  take_first: str | None = None) -> list[TimeSeries]:
    """`parse_parameter!` for interpolation-valued parameters (read.jl:199-326)."""
    static = _by_node(m, f"{node_type} / static")
    time = _by_node(m, f"{node_type} / time") if time_table else {}
    out = []
    for nid in node_ids:
        in_static, in_time = nid.value in static, nid.value in time
        if in_static and in_time:
            raise ValueError(f"Data for {nid} found in both Static and Time tables.")
        if in_static:
            val = _coalesce(static[nid.value][0].get(param), default)
            if val is None:
                raise ValueError(f"{nid}: missing value for {param}")
            out.append(TimeSeries.static(float(val), kind=kind))
        elif in_time:
            rows = time[nid.value]
            if take_first is not None:
                first = rows[0][take_first]
                rows = [r for r in rows if r[take_first] == first]
            u = [float(_coalesce(r.get(param), default)) for r in rows]
            t = [m.seconds_since(r["time"]) for r in rows]
            periodic = bool(cyclic and cyclic.get(nid.value, False))
            if len(t) == 1:
                # a single row is a constant series
                u, t = [u[0], u[0]], [t[0], t[0] + 1.0]
            out.append(TimeSeries(u, t, kind=kind, periodic=periodic))
        else:
            if is_complete:
                raise ValueError(f"Data for {nid} found in neither Static nor Time table.")
            out.append(TimeSeries.static(float(default), kind=kind))
    return out


def _control_states(m: ModelTables, node_type: str, node_ids: list[NodeID],
                    params: list[str]) -> dict[tuple[int, str], dict[str, float]]:
    """control_mapping: (node value, control_state) -> {param: value} (read.jl:328-427)."""
    name = f"{node_type} / static"
    out: dict[tuple[int, str], dict[str, float]] = {}
    if name not in m.tables:
        return out
    for r in m.table(name):
        cs = r.get("control_state")
        if cs is None:
            continue
        d = out.setdefault((r["node_id"], cs), {})
        for p in params:
            if r.get(p) is not None and p not in d:
                d[p] = float(r[p])
    return out


_M = 1 << 62
# (in_min, in_max, out_min, out_max) per node type: validation.jl:76-116
N_NEIGHBOR_BOUNDS_FLOW = {
    "Basin": (0, _M, 0, _M), "LinearResistance": (1, 1, 1, 1), "ManningResistance": (1, 1, 1, 1),
    "TabulatedRatingCurve": (1, 1, 1, 1), "LevelBoundary": (0, _M, 0, _M), "FlowBoundary": (0, 0, 1, 1),
    "Pump": (1, 1, 1, 1), "Outlet": (1, 1, 1, 1), "Terminal": (1, _M, 0, 0), "Junction": (1, _M, 1, _M),
    "PidControl": (0, 0, 0, 0), "ContinuousControl": (0, 0, 0, 0), "DiscreteControl": (0, 0, 0, 0),
    "UserDemand": (1, _M, 1, 1), "LevelDemand": (0, 0, 0, 0), "FlowDemand": (0, 0, 0, 0),
    "Observation": (0, 0, 0, 0),
}
N_NEIGHBOR_BOUNDS_CONTROL = {
    "Basin": (0, 1, 0, 0), "LinearResistance": (0, 1, 0, 0), "ManningResistance": (0, 1, 0, 0),
    "TabulatedRatingCurve": (0, 1, 0, 0), "LevelBoundary": (0, 0, 0, 0), "FlowBoundary": (0, 0, 0, 0),
    "Pump": (0, 2, 0, 0), "Outlet": (0, 2, 0, 0), "Terminal": (0, 0, 0, 0), "Junction": (0, 0, 0, 0),
    "PidControl": (0, 1, 1, 1), "ContinuousControl": (0, 0, 1, _M), "DiscreteControl": (0, 0, 1, _M),
    "UserDemand": (0, 0, 0, 0), "LevelDemand": (0, 0, 1, _M), "FlowDemand": (0, 0, 1, 1),
    "Observation": (0, 0, 0, 0),
}


def n_neighbor_errors(g: Graph) -> list[str]:
    """`valid_n_neighbors` (core/src/validation.jl:544-591): the messages of every violated bound,
    node by node in vertex order, flow links before control links."""
    out = []
    for code in sorted(g.label_of):
        nid = g.label_of[code]
        for link_type, bounds in (("flow", N_NEIGHBOR_BOUNDS_FLOW), ("control", N_NEIGHBOR_BOUNDS_CONTROL)):
            if nid.type not in bounds:
                raise ValueError(f"'n_neighbor_bounds_{link_type}' not defined for {nid.type}.")
            in_min, in_max, out_min, out_max = bounds[nid.type]
            n_in, n_out = len(g.inneighbors(nid, link_type)), len(g.outneighbors(nid, link_type))
            if n_in < in_min:
                out.append(f"{nid} must have at least {in_min} {link_type} inneighbor(s) (got {n_in}).")
            if n_in > in_max:
                out.append(f"{nid} can have at most {in_max} {link_type} inneighbor(s) (got {n_in}).")
            if n_out < out_min:
                out.append(f"{nid} must have at least {out_min} {link_type} outneighbor(s) (got {n_out}).")
            if n_out > out_max:
                out.append(f"{nid} can have at most {out_max} {link_type} outneighbor(s) (got {n_out}).")
    return out


def build_params(m: ModelTables) -> Params:
    g, node_ids_all, internal, external = create_graph(m)
    errs = n_neighbor_errors(g)
    if errs:
        raise ValueError("Invalid number of connections for certain node types. " + " ".join(errs))
    by_type: dict[str, list[NodeID]] = {}
    for nid in node_ids_all:
        by_type.setdefault(nid.type, []).append(nid)
    by_value = {n.value: n for n in node_ids_all}
    ids = lambda t: by_type.get(t, [])  # noqa: E731
    cyclic = {r["node_id"]: r["cyclic_time"] for r in m.node}
    depth_threshold = m.solver_option("depth_threshold")
    ldt = m.solver_option("level_difference_threshold")
    itp_cfg = {**INTERPOLATION_DEFAULTS, **m.interpolation}

    def inflow_link(nid):
        srcs = g.inflow_ids(nid)
        if len(srcs) != 1:
            raise ValueError(f"{nid} must have exactly one flow inneighbour, got {len(srcs)}")
        return g.edge[(srcs[0], nid)]

    def outflow_link(nid):
        dsts = g.outflow_ids(nid)
        if len(dsts) != 1:
            raise ValueError(f"{nid} must have exactly one flow outneighbour, got {len(dsts)}")
        return g.edge[(nid, dsts[0])]

    # ------------------------------------------------------------------ Basin
    b_ids = ids("Basin")
    if not b_ids:
        raise ValueError("Models without states are unsupported, please add a Basin node.")
    prof = _group(m.table("Basin / profile"), "node_id") if "Basin / profile" in m.tables else {}
    levels, areas, dSdhs, slopes, cumSs, aslopes = [], [], [], [], [], []
    l2a_tables: list[LinearTable] = []
    for nid in b_ids:
        rows = prof.get(nid.value)
        if not rows or len(rows) < 2:
            raise ValueError(f"{nid} needs at least two profile rows")
        lv = [float(r["level"]) for r in rows]
        ar = [r["area"] for r in rows]
        st = [r["storage"] for r in rows]
        if all(s is None for s in st):
            if any(a is None for a in ar):
                raise ValueError(f"{nid}: profile needs area or storage")
            st = trapz_integrate([float(a) for a in ar], lv)
        else:
            st = [float(s) for s in st]
        if ar[0] is None:
            dS = finite_difference(st, lv)
        else:
            dS = finite_difference(st, lv, float(ar[0]))
        for j in range(len(dS) - 1):
            if dS[j + 1] < 0:
                raise ValueError(
                    f"Invalid profile for {nid}. The step from (h={_jl(lv[j])}, S={_jl(st[j])}) to "
                    f"(h={_jl(lv[j + 1])}, S={_jl(st[j + 1])}) implies a decreasing area compared to "
                    "lower points in the profile, which is not allowed.")
        # storage_to_level = invert_integral(LinearInterpolation(dS_dh, level))
        inv = LinearTable(dS, lv, left="extension", right="constant")
        if not all(a is None for a in ar):
            l2a = LinearTable([float(a) for a in ar], lv, left="constant", right="constant")
        else:
            l2a = inv
        levels.append(lv)
        dSdhs.append(list(inv.u))
        slopes.append(list(inv.slope))
        cumSs.append(list(inv.I))
        areas.append(list(l2a.u))
        aslopes.append(list(l2a.slope))
        l2a_tables.append(l2a)

    forcing_names = ("precipitation", "surface_runoff", "potential_evaporation", "drainage",
                     "infiltration")
    forcing = {
        p: _parse_series(m, "Basin", b_ids, p, default=math.nan, is_complete=False, cyclic=cyclic)
        for p in forcing_names
    }
    vertical_flux = {p: np.zeros(len(b_ids)) for p in forcing_names}
    for p in forcing_names:  # update_basin!(basin, 0.0), callback.jl:855-866
        for i, s in enumerate(forcing[p]):
            v = s(0.0)
            if not math.isnan(v):
                vertical_flux[p][i] = v

    state_rows = {r["node_id"]: r for r in m.table("Basin / state")} \
        if "Basin / state" in m.tables else {}
    storage0 = np.zeros(len(b_ids))
    low_thr = np.zeros(len(b_ids))
    for i, nid in enumerate(b_ids):
        if nid.value not in state_rows:
            raise ValueError(f"Unexpected 'Basin / state' length: no state for {nid}")
        lvl = float(state_rows[nid.value]["level"])
        bottom = levels[i][0]
        if lvl < bottom:
            raise ValueError(f"The initial level ({lvl}) of {nid} is below the bottom ({bottom}).")
        # get_storage_from_level (util.jl:2-9)
        storage0[i] = l2a_tables[i].integral_from_start(lvl)
        low_thr[i] = l2a_tables[i].integral_from_start(bottom + depth_threshold)

    basin = BasinP(
        node_id=b_ids,
        inflow_ids=[g.inflow_ids(n) for n in b_ids],
        outflow_ids=[g.outflow_ids(n) for n in b_ids],
        storage0=storage0,
        low_storage_threshold=low_thr,
        levels=levels, areas=areas, dSdh=dSdhs, dSdh_slope=slopes, cumS=cumSs,
        area_slope=aslopes, forcing=forcing, vertical_flux=vertical_flux,
    )

    def basin_bottom(nid: NodeID) -> float:
        return levels[nid.idx - 1][0]

    nodes: dict[str, dict[str, Any]] = {}

    def control_type_of(nid: NodeID) -> int:
        ctrl = [c for c in g.inneighbors(nid, "control") if c.type != "FlowDemand"]
        if len(ctrl) > 1:
            raise ValueError(f"{nid} has more than 1 control inneighbors.")
        if len(ctrl) == 1:
            return {"PidControl": CONTROL_PID,
                    "ContinuousControl": CONTROL_CONTINUOUS}.get(ctrl[0].type, CONTROL_NONE)
        return CONTROL_NONE

    def flow_demand_of(nid: NodeID) -> NodeID | None:
        for c in g.inneighbors(nid, "control"):
            if c.type == "FlowDemand":
                return c
        return None

    # ------------------------------------------------------------------ LinearResistance
    lr_ids = ids("LinearResistance")
    nodes["linear_resistance"] = dict(
        node_id=lr_ids,
        inflow_link=[inflow_link(n) for n in lr_ids],
        outflow_link=[outflow_link(n) for n in lr_ids],
        # parse_parameter! on a Number takes the last row of the node's group (read.jl:66)
        resistance=as_f64([_last_static(m, "LinearResistance", n, "resistance") for n in lr_ids]),
        max_flow_rate=as_f64([
            _coalesce(_last_static(m, "LinearResistance", n, "max_flow_rate"), math.inf)
            for n in lr_ids]),
        control_mapping=_control_states(m, "LinearResistance", lr_ids,
                                        ["resistance", "max_flow_rate"]),
    )

    # ------------------------------------------------------------------ ManningResistance
    mn_ids = ids("ManningResistance")
    mn = dict(node_id=mn_ids,
              inflow_link=[inflow_link(n) for n in mn_ids],
              outflow_link=[outflow_link(n) for n in mn_ids])
    for p in ("length", "manning_n", "profile_width", "profile_slope"):
        mn[p] = as_f64([_last_static(m, "ManningResistance", n, p) for n in mn_ids])
    for n in mn_ids:
        for end in (g.inflow_ids(n)[0], g.outflow_ids(n)[0]):
            if end.type != "Basin":
                raise ValueError(f"{n} must connect two Basins")
    mn["upstream_bottom"] = as_f64([basin_bottom(g.inflow_ids(n)[0]) for n in mn_ids])
    mn["downstream_bottom"] = as_f64([basin_bottom(g.outflow_ids(n)[0]) for n in mn_ids])
    mn["control_mapping"] = _control_states(
        m, "ManningResistance", mn_ids, ["length", "manning_n", "profile_width", "profile_slope"])
    nodes["manning_resistance"] = mn

    # ------------------------------------------------------------------ TabulatedRatingCurve
    trc_ids = ids("TabulatedRatingCurve")
    trc_static = _group(m.table("TabulatedRatingCurve / static"), "node_id") \
        if "TabulatedRatingCurve / static" in m.tables else {}
    trc_time = _group(m.table("TabulatedRatingCurve / time"), "node_id") \
        if "TabulatedRatingCurve / time" in m.tables else {}
    interpolations: list[PchipTable] = []
    current_index: list[IndexLookup] = []
    trc_cm: dict[tuple[int, str], dict[str, float]] = {}
    for nid in trc_ids:
        in_static, in_time = nid.value in trc_static, nid.value in trc_time
        if in_static and in_time:
            raise ValueError(f"Data for {nid} found in both Static and Time tables.")
        if not in_static and not in_time:
            raise ValueError(f"Data for {nid} found in neither Static nor Time table.")
        if in_static:
            # groups by control_state in sorted order (missing last)
            groups: dict[Any, list[dict]] = {}
            for r in trc_static[nid.value]:
                groups.setdefault(r.get("control_state"), []).append(r)
            last_index = 0
            for cs, rows in groups.items():
                interpolations.append(qh_interpolation(
                    nid, [r["level"] for r in rows], [r["flow_rate"] for r in rows]))
                last_index = len(interpolations)
                if cs is not None:
                    trc_cm[(nid.value, cs)] = {"table": float(last_index)}
            current_index.append(IndexLookup.static(last_index))
        else:
            groups = {}
            for r in trc_time[nid.value]:
                groups.setdefault(m.seconds_since(r["time"]), []).append(r)
            lookup_t, lookup_i = [], []
            for t_, rows in groups.items():
                interpolations.append(qh_interpolation(
                    nid, [r["level"] for r in rows], [r["flow_rate"] for r in rows]))
                lookup_i.append(len(interpolations))
                lookup_t.append(t_)
            per = bool(cyclic.get(nid.value, False))
            if per:
                lookup_i[-1] = lookup_i[0]
            if len(lookup_t) == 1:
                lookup_t, lookup_i = [lookup_t[0], lookup_t[0] + 1.0], [lookup_i[0]] * 2
            current_index.append(IndexLookup(lookup_i, lookup_t, periodic=per))
    # valid_tabulated_curve_level (validation.jl:474-499): over the whole time series the lowest
    # level of a curve may not lie below the bottom of the upstream Basin
    for nid, lookup in zip(trc_ids, current_index):
        src = g.inflow_ids(nid)[0]
        if src.type != "Basin":
            continue
        for k in lookup.u:
            h_min = interpolations[int(k) - 1].t[1]
            if h_min < basin_bottom(src):
                raise ValueError(f"Lowest level of {nid} is lower than bottom of upstream {src}")
    nodes["tabulated_rating_curve"] = dict(
        node_id=trc_ids,
        inflow_link=[inflow_link(n) for n in trc_ids],
        outflow_link=[outflow_link(n) for n in trc_ids],
        max_downstream_level=_parse_series(m, "TabulatedRatingCurve", trc_ids,
                                           "max_downstream_level", default=math.inf,
                                           cyclic=cyclic),
        interpolations=interpolations,
        current_interpolation_index=current_index,
        control_mapping=trc_cm,
    )

    # ------------------------------------------------------------------ boundaries
    lb_ids = ids("LevelBoundary")
    nodes["level_boundary"] = dict(
        node_id=lb_ids,
        level=_parse_series(m, "LevelBoundary", lb_ids, "level", cyclic=cyclic),
    )
    fb_ids = ids("FlowBoundary")
    fb_kind = {"block": "constant", "linear": "linear"}[itp_cfg["flow_boundary"]]
    fb_rate = _parse_series(m, "FlowBoundary", fb_ids, "flow_rate", kind=fb_kind, cyclic=cyclic)
    for s in fb_rate:
        if any(u < 0 for u in s.u):
            raise ValueError("Currently negative flow rates for FlowBoundary are not supported.")
    nodes["flow_boundary"] = dict(
        node_id=fb_ids,
        outflow_link=[outflow_link(n) for n in fb_ids],
        flow_rate=fb_rate,
    )

    # ------------------------------------------------------------------ Pump / Outlet
    for node_type, key in (("Pump", "pump"), ("Outlet", "outlet")):
        p_ids = ids(node_type)
        static = _group(m.table(f"{node_type} / static"), "node_id") \
            if f"{node_type} / static" in m.tables else {}
        time = _group(m.table(f"{node_type} / time"), "node_id") \
            if f"{node_type} / time" in m.tables else {}
        flow_rate = np.zeros(len(p_ids))
        td_flow_rate: list[TimeSeries | None] = []
        for i, nid in enumerate(p_ids):
            if nid.value in static and nid.value in time:
                raise ValueError(f"Data for {nid} found in both Static and Time tables.")
            if nid.value in static:
                flow_rate[i] = float(static[nid.value][-1]["flow_rate"])
                td_flow_rate.append(None)
            elif nid.value in time:
                rows = time[nid.value]
                u = [float(r["flow_rate"]) for r in rows]
                t = [m.seconds_since(r["time"]) for r in rows]
                if len(t) == 1:
                    u, t = [u[0]] * 2, [t[0], t[0] + 1.0]
                td_flow_rate.append(TimeSeries(u, t, periodic=bool(cyclic.get(nid.value))))
            else:
                raise ValueError(f"Data for {nid} found in neither Static nor Time table.")
        d = dict(
            node_id=p_ids,
            inflow_link=[inflow_link(n) for n in p_ids],
            outflow_link=[outflow_link(n) for n in p_ids],
            flow_rate=flow_rate,
            time_dependent_flow_rate=td_flow_rate,
            min_flow_rate=_parse_series(m, node_type, p_ids, "min_flow_rate", default=0.0,
                                        cyclic=cyclic),
            max_flow_rate=_parse_series(m, node_type, p_ids, "max_flow_rate", default=math.inf,
                                        cyclic=cyclic),
            min_upstream_level=_parse_series(m, node_type, p_ids, "min_upstream_level",
                                             default=-math.inf, cyclic=cyclic),
            max_downstream_level=_parse_series(m, node_type, p_ids, "max_downstream_level",
                                               default=math.inf, cyclic=cyclic),
            control_type=[control_type_of(n) for n in p_ids],
            flow_demand_id=[flow_demand_of(n) for n in p_ids],
            control_mapping=_control_states(
                m, node_type, p_ids,
                ["flow_rate", "min_flow_rate", "max_flow_rate", "min_upstream_level",
                 "max_downstream_level"]),
        )
        # valid_flow_rates (validation.jl:321-372): no negative flow rates, in the control states
        # or — for nodes without control states — in the static / time values
        controlled = {k[0] for k in d["control_mapping"]}
        for (nv, cs), upd in d["control_mapping"].items():
            if upd.get("flow_rate", 0.0) < 0.0:
                raise ValueError(f"Negative flow rate(s) found. node_id = {node_type} #{nv} control_state = {cs}")
        for i, nid in enumerate(p_ids):
            if nid.value in controlled:
                continue
            ts = td_flow_rate[i]
            if (ts is not None and min(ts.u) < 0.0) or (ts is None and flow_rate[i] < 0.0):
                raise ValueError(f"Negative flow rate(s) for {nid} found.")
        # valid_min_upstream_level! (validation.jl:453-472): default -Inf -> basin bottom
        for i, nid in enumerate(p_ids):
            src = d["inflow_link"][i].link[0]
            if src.type == "Basin":
                s = d["min_upstream_level"][i]
                bottom = basin_bottom(src)
                if all(u == -math.inf for u in s.u):
                    s.u = [bottom for _ in s.u]
                elif min(s.u) < bottom:
                    raise ValueError(
                        f"Minimum upstream level of {nid} is lower than bottom of upstream {src}")
        nodes[key] = d

    # ------------------------------------------------------------------ demand priorities
    prios: set[int] = set()
    for name in ("UserDemand / static", "UserDemand / time", "FlowDemand / static",
                 "FlowDemand / time"):
        for r in m.tables.get(name, []):
            prios.add(int(_coalesce(r.get("demand_priority"), 0)))
    demand_priorities = sorted(prios)

    def parse_demand(node_type: str, d_ids: list[NodeID]):
        static = _group(m.table(f"{node_type} / static"), "node_id") \
            if f"{node_type} / static" in m.tables else {}
        time = _group(m.table(f"{node_type} / time"), "node_id") \
            if f"{node_type} / time" in m.tables else {}
        npr = len(demand_priorities)
        has = np.zeros((len(d_ids), npr), dtype=bool)
        itps: list[list[TimeSeries | None]] = [[None] * npr for _ in d_ids]
        demand = np.zeros((len(d_ids), npr))
        from_ts = [False] * len(d_ids)
        for i, nid in enumerate(d_ids):
            if nid.value in static and nid.value in time:
                raise ValueError(f"Data for {nid} found in both Static and Time tables.")
            if nid.value in static:
                for r in static[nid.value]:
                    k = demand_priorities.index(int(_coalesce(r.get("demand_priority"), 0)))
                    has[i, k] = True
                    v = float(_coalesce(r.get("demand"), 0.0))
                    itps[i][k] = TimeSeries([v, v], [0.0, 1.0])
                    demand[i, k] = v
            elif nid.value in time:
                from_ts[i] = True
                gp: dict[int, list[dict]] = {}
                for r in time[nid.value]:
                    gp.setdefault(int(_coalesce(r.get("demand_priority"), 0)), []).append(r)
                for pr, rows in gp.items():
                    k = demand_priorities.index(pr)
                    has[i, k] = True
                    u = [float(_coalesce(r["demand"], 0.0)) for r in rows]
                    t = [m.seconds_since(r["time"]) for r in rows]
                    if len(t) == 1:
                        u, t = [u[0]] * 2, [t[0], t[0] + 1.0]
                    itps[i][k] = TimeSeries(u, t, periodic=bool(cyclic.get(nid.value)))
                    demand[i, k] = itps[i][k](0.0)
            else:
                raise ValueError(f"Data for {nid} found in neither Static nor Time table.")
        return has, itps, demand, from_ts

    # ------------------------------------------------------------------ UserDemand
    ud_ids = ids("UserDemand")
    has, itps, demand, from_ts = parse_demand("UserDemand", ud_ids)
    inflow_links = [[g.edge[(s, n)] for s in g.inflow_ids(n)] for n in ud_ids]
    offsets = [0]
    for links in inflow_links:
        offsets.append(offsets[-1] + len(links))
    min_level = []
    for nid in ud_ids:
        rows = _by_node(m, "UserDemand / static").get(nid.value) or \
            _by_node(m, "UserDemand / time").get(nid.value)
        min_level.append(float(rows[-1]["min_level"]))
    allocated = np.full((len(ud_ids), len(demand_priorities)), math.inf)
    allocated[~has] = 0.0
    nodes["user_demand"] = dict(
        node_id=ud_ids,
        demand_priorities=demand_priorities,
        inflow_links=inflow_links,
        inflow_link_offsets=offsets,
        inflow_link_allocated=[[math.inf] * len(l) for l in inflow_links],
        outflow_link=[outflow_link(n) for n in ud_ids],
        has_demand_priority=has,
        demand=demand,
        demand_interpolation=itps,
        demand_from_timeseries=from_ts,
        allocated=allocated,
        return_factor=_parse_series(m, "UserDemand", ud_ids, "return_factor", cyclic=cyclic,
                                    take_first="demand_priority"),
        min_level=as_f64(min_level),
    )

    # ------------------------------------------------------------------ FlowDemand
    fd_ids = ids("FlowDemand")
    fhas, fitps, fdemand, _ = parse_demand("FlowDemand", fd_ids)
    nodes["flow_demand"] = dict(node_id=fd_ids, has_demand_priority=fhas,
                                demand_interpolation=fitps)

    # ------------------------------------------------------------------ state layout
    ud_in_ids = [ud_ids[i] for i in range(len(ud_ids)) for _ in inflow_links[i]]
    pid_ids = ids("PidControl")
    u_ids = {
        "tabulated_rating_curve": trc_ids,
        "pump": ids("Pump"),
        "outlet": ids("Outlet"),
        "user_demand_inflow": ud_in_ids,
        "user_demand_outflow": ud_ids,
        "linear_resistance": lr_ids,
        "manning_resistance": mn_ids,
        "evaporation": b_ids,
        "infiltration": b_ids,
        "integral": pid_ids,
    }
    state_node_id: list[NodeID] = []
    state_ranges: dict[str, tuple[int, int]] = {}
    for comp in STATE_COMPONENTS:
        start = len(state_node_id)
        state_node_id.extend(u_ids[comp])
        state_ranges[comp] = (start, len(state_node_id))

    # get_state_flow_links (util.jl:798-854)
    s_in: list[LinkMeta | None] = []
    s_out: list[LinkMeta | None] = []
    for comp in STATE_COMPONENTS:
        if comp in ("tabulated_rating_curve", "pump", "outlet", "linear_resistance",
                    "manning_resistance"):
            for k in range(len(nodes[comp]["node_id"])):
                s_in.append(nodes[comp]["inflow_link"][k])
                s_out.append(nodes[comp]["outflow_link"][k])
        elif comp == "user_demand_inflow":
            for links in inflow_links:
                for lm in links:
                    s_in.append(lm)
                    s_out.append(None)
        elif comp == "user_demand_outflow":
            for k in range(len(ud_ids)):
                s_in.append(None)
                s_out.append(nodes["user_demand"]["outflow_link"][k])
    link_to_state_idx = {lm.link: i for i, lm in enumerate(s_in) if lm is not None}

    # ------------------------------------------------------------------ PidControl
    pid_static = _group(m.table("PidControl / static"), "node_id") \
        if "PidControl / static" in m.tables else {}
    pid_time = _group(m.table("PidControl / time"), "node_id") \
        if "PidControl / time" in m.tables else {}
    listen = []
    for nid in pid_ids:
        rows = pid_static.get(nid.value) or pid_time.get(nid.value)
        if not rows:
            raise ValueError(f"Data for {nid} found in neither Static nor Time table.")
        ln = by_value[int(rows[-1]["listen_node_id"])]
        if ln.type != "Basin":
            raise ValueError(f"Listen node {ln} of {nid} is not a Basin")
        listen.append(ln)
    pid_target_nodes = []
    for nid in pid_ids:
        tgt = g.outneighbors(nid, "control")
        if len(tgt) != 1 or tgt[0].type not in ("Pump", "Outlet"):
            raise ValueError(f"{nid} must control exactly one Pump or Outlet")
        # valid_pid_connectivity (validation.jl:373-400)
        ln = listen[len(pid_target_nodes)]
        if ln not in (g.inflow_ids(tgt[0])[0], g.outflow_ids(tgt[0])[0]):
            raise ValueError(f"PID listened {ln} is not on either side of controlled {tgt[0]}.")
        pid_target_nodes.append(tgt[0])
    nodes["pid_control"] = dict(
        node_id=pid_ids,
        listen_node_id=listen,
        target=_parse_series(m, "PidControl", pid_ids, "target", cyclic=cyclic),
        proportional=_parse_series(m, "PidControl", pid_ids, "proportional", cyclic=cyclic),
        integral=_parse_series(m, "PidControl", pid_ids, "integral", cyclic=cyclic),
        derivative=_parse_series(m, "PidControl", pid_ids, "derivative", cyclic=cyclic),
        target_node=pid_target_nodes,
        control_mapping=_control_states(m, "PidControl", pid_ids,
                                        ["target", "proportional", "integral", "derivative"]),
    )

    # ------------------------------------------------------------------ compound variables
    def subvariable(row) -> dict[str, Any]:
        ln = by_value[int(row["listen_node_id"])]
        if ln.type == "Junction":
            raise ValueError("Cannot listen to Junction node")
        var = row["variable"]
        sv = dict(listen_node_id=ln, variable=var, weight=float(_coalesce(row.get("weight"), 1.0)),
                  look_ahead=float(_coalesce(row.get("look_ahead"), 0.0)))
        # get_cache_ref (util.jl:491-531)
        if ln.type == "Basin" and var == "level":
            sv.update(kind=SV_BASIN_LEVEL, idx=ln.idx - 1)
        elif ln.type == "Basin" and var == "storage":
            sv.update(kind=SV_BASIN_STORAGE, idx=ln.idx - 1)
        elif var == "flow_rate" and ln.type != "FlowBoundary":
            if ln.type not in CONSERVATIVE:
                raise ValueError(f"Cannot listen to flow_rate of {ln}")
            comp = {"Pump": "pump", "Outlet": "outlet",
                    "TabulatedRatingCurve": "tabulated_rating_curve",
                    "LinearResistance": "linear_resistance",
                    "ManningResistance": "manning_resistance"}[ln.type]
            sv.update(kind=SV_DU_STATE, idx=state_ranges[comp][0] + ln.idx - 1)
        elif var == "level" and ln.type == "LevelBoundary":
            sv.update(kind=SV_CONSTANT, idx=0,
                      series=nodes["level_boundary"]["level"][ln.idx - 1])
        elif var == "flow_rate" and ln.type == "FlowBoundary":
            sv.update(kind=SV_CONSTANT, idx=0,
                      series=nodes["flow_boundary"]["flow_rate"][ln.idx - 1])
        else:
            raise ValueError(f"Unsupported condition variable {var} on {ln}.")
        return sv

    # ------------------------------------------------------------------ ContinuousControl
    cc_ids = ids("ContinuousControl")
    cc_var = _group(m.table("ContinuousControl / variable"), "node_id") \
        if "ContinuousControl / variable" in m.tables else {}
    cc_fun = _group(m.table("ContinuousControl / function"), "node_id") \
        if "ContinuousControl / function" in m.tables else {}
    cc_funcs, cc_cvars, cc_targets = [], [], []
    for nid in cc_ids:
        rows = cc_fun.get(nid.value, [])
        if len(rows) < 2:
            raise ValueError("There must be at least 2 data points in a ContinuousControl function.")
        cv = {r["controlled_variable"] for r in rows}
        if len(cv) != 1:
            raise ValueError("There must be a unique 'controlled_variable'.")
        inp = [float(r["input"]) for r in rows]
        outp = [float(r["output"]) for r in rows]
        if len(rows) == 2:  # read.jl:1181-1186
            inp.insert(1, sum(inp) / 2)
            outp.insert(1, sum(outp) / 2)
        cc_funcs.append(PchipTable(outp, inp, left="linear", right="linear"))
        cc_cvars.append([subvariable(r) for r in cc_var.get(nid.value, [])])
        tgt = g.outneighbors(nid, "control")
        if len(tgt) != 1 or tgt[0].type not in ("Pump", "Outlet") or cv != {"flow_rate"}:
            raise ValueError(f"{nid}: only flow_rate of one Pump/Outlet can be controlled")
        cc_targets.append(tgt[0])
    nodes["continuous_control"] = dict(node_id=cc_ids, func=cc_funcs, compound_variable=cc_cvars,
                                       target_node=cc_targets)

    # ------------------------------------------------------------------ DiscreteControl
    dc_ids = ids("DiscreteControl")
    dc_var = _group(m.table("DiscreteControl / variable"), "node_id") \
        if "DiscreteControl / variable" in m.tables else {}
    dc_cond = _group(m.table("DiscreteControl / condition"), "node_id") \
        if "DiscreteControl / condition" in m.tables else {}
    dc_logic = _group(m.table("DiscreteControl / logic"), "node_id") \
        if "DiscreteControl / logic" in m.tables else {}
    dc_cvars: list[list[dict[str, Any]]] = []
    dc_logic_maps: list[dict[tuple[bool, ...], str]] = []
    for nid in dc_ids:
        cvars = []
        vars_by_cv = _group(dc_var.get(nid.value, []), "compound_variable_id")
        conds_by_cv: dict[int, list[dict]] = {}
        for r in dc_cond.get(nid.value, []):
            conds_by_cv.setdefault(r["compound_variable_id"], []).append(r)
        for cv_id, conds in conds_by_cv.items():
            if cv_id not in vars_by_cv:
                raise ValueError(
                    f"compound_variable_id {cv_id} for {nid} in condition table but not in "
                    "variable table")
            thr_hi, thr_lo = [], []
            for cid, rows in _group(conds, "condition_id").items():
                ts = [0.0 if r["time"] is None else m.seconds_since(r["time"]) for r in rows]
                if len(set(ts)) != len(ts):
                    raise ValueError(f"Condition {cid} for {nid} has repeated timestamps")
                hi = [float(r["threshold_high"]) for r in rows]
                lo = [float(_coalesce(r.get("threshold_low"), r["threshold_high"])) for r in rows]
                if len(ts) == 1:
                    ts, hi, lo = [ts[0], ts[0] + 1.0], hi * 2, lo * 2
                per = bool(cyclic.get(nid.value))
                thr_hi.append(TimeSeries(hi, ts, periodic=per))
                thr_lo.append(TimeSeries(lo, ts, periodic=per))
            cvars.append(dict(subvariables=[subvariable(r) for r in vars_by_cv[cv_id]],
                              threshold_high=thr_hi, threshold_low=thr_lo))
        if not cvars:
            raise ValueError(f"Missing data for {nid}.")
        dc_cvars.append(cvars)
        lm = {r["truth_state"]: r["control_state"] for r in dc_logic.get(nid.value, [])}
        dc_logic_maps.append(expand_logic_mapping(lm, nid))
    nodes["discrete_control"] = dict(
        node_id=dc_ids,
        compound_variables=dc_cvars,
        logic_mapping=dc_logic_maps,
        controlled_nodes=[g.outneighbors(n, "control") for n in dc_ids],
    )
    nodes["terminal"] = dict(node_id=ids("Terminal"))
    errs = discrete_control_errors(nodes, g, m.t_end)
    if errs:
        raise ValueError("Invalid discrete control state definition(s). " + " ".join(errs))

    p = Params(
        tables=m, graph=g, node_ids_all=node_ids_all, basin=basin, nodes=nodes,
        state_node_id=state_node_id, state_ranges=state_ranges, state_inflow_link=s_in,
        state_outflow_link=s_out, link_to_state_idx=link_to_state_idx,
        internal_flow_links=internal, external_flow_links=external,
        level_difference_threshold=ldt,
    )
    p.tstops = timeseries_tstops(p, m.t_end)
    return p


def _last_static(m: ModelTables, node_type: str, nid: NodeID, param: str):
    rows = _by_node(m, f"{node_type} / static").get(nid.value)
    if not rows:
        raise ValueError(f"Data for {nid} found in neither Static nor Time table.")
    return rows[-1].get(param)


def qh_interpolation(nid: NodeID, level, flow_rate) -> PchipTable:
    """core/src/util.jl:101-141."""
    level = [float(x) for x in level]
    flow_rate = [float(x) for x in flow_rate]
    if len(level) < 2:
        raise ValueError(f"{nid}: At least two datapoints are needed.")
    if flow_rate[0] != 0.0:
        raise ValueError(f"{nid}: The `flow_rate` must start at 0.")
    if len(set(level)) != len(level):
        raise ValueError(f"{nid}: The `level` cannot be repeated.")
    if any(b - a < 0 for a, b in zip(flow_rate, flow_rate[1:])):
        raise ValueError(f"{nid}: The `flow_rate` cannot decrease with increasing `level`.")
    level = [level[0] - 1.0] + level
    flow_rate = [0.0] + flow_rate
    return PchipTable(flow_rate, level, left="constant", right="linear")


_NODE_KEY = {"Pump": "pump", "Outlet": "outlet", "LinearResistance": "linear_resistance",
             "ManningResistance": "manning_resistance", "TabulatedRatingCurve": "tabulated_rating_curve",
             "PidControl": "pid_control"}


def discrete_control_errors(nodes: dict[str, dict[str, Any]], g: Graph, t_end: float) -> list[str]:
    """`valid_discrete_control` (core/src/validation.jl:617-718), the messages in its order."""
    out = []
    dc = nodes["discrete_control"]
    for nid, cvars, lm, targets in zip(dc["node_id"], dc["compound_variables"], dc["logic_mapping"],
                                       dc["controlled_nodes"]):
        n_cond = sum(len(cv["threshold_high"]) for cv in cvars)
        states = list(dict.fromkeys(lm.values()))
        wrong = ["".join("T" if b else "F" for b in ts) for ts in lm if len(ts) != n_cond]
        if wrong:
            lst = ", ".join(f'"{w}"' for w in wrong)
            out.append(f"{nid} has {n_cond} condition(s), which is inconsistent with these truth state(s): [{lst}].")
        for tgt in targets:
            cm = nodes[_NODE_KEY[tgt.type]].get("control_mapping", {}) if tgt.type in _NODE_KEY else {}
            defined = {cs for (nv, cs) in cm if nv == tgt.value}
            undefined = [s for s in states if s not in defined]
            if undefined:
                lst = ", ".join(f'"{s}"' for s in undefined)
                out.append(f"These control states from {nid} are not defined for controlled {tgt}: [{lst}].")
        for cv in cvars:
            for hi, lo in zip(cv["threshold_high"], cv["threshold_low"]):
                if any(a > b for a, b in zip(lo.u, hi.u)):
                    out.append("threshold_low is not less than or equal to threshold_high for "
                               f"'{nid}'")
        for cv in cvars:
            for sv in cv["subvariables"]:
                la, ln, var = sv["look_ahead"], sv["listen_node_id"], sv["variable"]
                if la == 0:
                    continue
                if ln.type not in ("FlowBoundary", "LevelBoundary"):
                    out.append(f"Look ahead supplied for non-timeseries listen variable '{var}' "
                               f"from listen node {ln}.")
                elif la < 0:
                    out.append(f"Negative look ahead supplied for listen variable '{var}' "
                               f"from listen node {ln}.")
                elif t_end + la > sv["series"].t[-1]:
                    out.append(f"Look ahead for listen variable '{var}' from listen node {ln} "
                               "goes past timeseries end during simulation.")
    return out


def expand_logic_mapping(logic: dict[str, str], nid: NodeID) -> dict[tuple[bool, ...], str]:
    """core/src/util.jl:215-268."""
    import itertools
    import re

    out: dict[tuple[bool, ...], str] = {}
    for truth_state, control_state in logic.items():
        if not re.fullmatch(r"[TF\*]+", truth_state):
            raise ValueError(f"Truth state '{truth_state}' contains illegal characters or is empty.")
        nw = truth_state.count("*")
        for sub in itertools.product([True, False], repeat=nw):
            k = 0
            ts = []
            for ch in truth_state:
                if ch == "*":
                    ts.append(sub[k])
                    k += 1
                else:
                    ts.append(ch == "T")
            key = tuple(ts)
            if key in out and out[key] != control_state:
                tf = "".join("T" if b else "F" for b in key)
                raise AssertionError(
                    f"Multiple control states found for {nid} for truth state `{tf}`: "
                    f"{sorted([control_state, out[key]])}.")
            out[key] = control_state
    return out


def timeseries_tstops(p: Params, t_end: float) -> list[float]:
    """core/src/util.jl:1009-1067 + saveat multiples (model.jl:194-196)."""
    ts: set[float] = set()
    n = p.nodes

    def add(series):
        if series is not None:
            ts.update(series.tstops(t_end))

    for s in p.basin.forcing["precipitation"]:
        add(s)
    for key, fields in (
        ("flow_boundary", ["flow_rate"]),
        ("level_boundary", ["level"]),
        ("pid_control", ["target"]),
        ("tabulated_rating_curve", ["current_interpolation_index"]),
        ("user_demand", ["return_factor"]),
        ("pump", ["time_dependent_flow_rate", "min_flow_rate", "max_flow_rate",
                  "min_upstream_level", "max_downstream_level"]),
        ("outlet", ["time_dependent_flow_rate", "min_flow_rate", "max_flow_rate",
                    "min_upstream_level", "max_downstream_level"]),
    ):
        for f in fields:
            for s in n[key][f]:
                add(s)
    for key in ("user_demand", "flow_demand"):
        for row in n[key]["demand_interpolation"]:
            for s in row:
                add(s)
    for cvars in n["discrete_control"]["compound_variables"]:
        for cv in cvars:
            for s in cv["threshold_high"]:
                add(s)
    saveat = p.tables.solver_option("saveat")
    if saveat and math.isfinite(saveat) and saveat > 0:
        k = 0
        while k * saveat <= t_end:
            ts.add(k * saveat)
            k += 1
    ts.add(0.0)
    ts.add(t_end)
    return sorted(ts)
