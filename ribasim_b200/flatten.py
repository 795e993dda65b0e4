"""Flatten `Params` into the index/value arrays of `rbs_network_desc` (include/ribasim_b200.h).

Everything here is integer layout or init-time table construction:

* CSR incidence `A` (basin -> signed state list) in the summation order of `reduce_state!`
  (core/src/util.jl:745-791);
* per-connector endpoints (`get_level` kinds, core/src/util.jl:170-191);
* the structural pattern of `J_intermediate = dg/du_reduced` (what the reference obtains
  from `TracerSparsityDetector`, core/src/differentiation.jl:396-430; golden rows/cols in
  core/test/utils_test.jl:266-271, 287-292);
* the pattern of `W_inner = A*J_intermediate - I/gamma` and, for every non-zero, the ordered
  list of `J_intermediate` entries it sums (`calc_J_inner!`, differentiation.jl:153-262).
"""

from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Any

import numpy as np

from .interp import as_f64, as_i32
from .network import (
    CONTROL_CONTINUOUS,
    CONTROL_NONE,
    CONTROL_PID,
    CT_LR,
    CT_MANNING,
    CT_OUTLET,
    CT_PUMP,
    CT_TRC,
    CT_UD_IN,
    CT_UD_OUT,
    KIND_BASIN,
    KIND_NONE,
    SV_BASIN_LEVEL,
    SV_BASIN_STORAGE,
    SV_CONSTANT,
    SV_DU_STATE,
    NodeID,
    Params,
    _endpoint,
)


@dataclass
class Flat:
    """Flat network: `ints` (scalars), `arrays` (numpy), ready for the C descriptor."""

    ints: dict[str, int] = field(default_factory=dict)
    floats: dict[str, float] = field(default_factory=dict)
    arrays: dict[str, np.ndarray] = field(default_factory=dict)
    # J_intermediate pattern in CSR (row = state) and CSC order (for golden tests)
    jac_rows_csc: np.ndarray | None = None
    jac_cols_csc: np.ndarray | None = None
    dc_state_names: list | None = None

    def __getitem__(self, k: str):
        if k in self.arrays:
            return self.arrays[k]
        if k in self.ints:
            return self.ints[k]
        return self.floats[k]


def _ragged(lists, dtype=np.float64):
    ptr = [0]
    flat: list[Any] = []
    for l in lists:
        flat.extend(l)
        ptr.append(len(flat))
    return as_i32(ptr), np.ascontiguousarray(np.asarray(flat, dtype=dtype))


def flatten(p: Params) -> Flat:
    f = Flat()
    n = p.nodes
    basin = p.basin
    n_b = len(basin.node_id)
    sr = p.state_ranges
    n_state = p.n_states
    n_conn = sr["evaporation"][0]
    n_pid = sr["integral"][1] - sr["integral"][0]
    f.ints.update(
        n_basin=n_b, n_pid=n_pid, n_state=n_state, n_conn=n_conn,
        n_reduced=n_b + n_pid,
        n_trc=len(n["tabulated_rating_curve"]["node_id"]),
        n_pump=len(n["pump"]["node_id"]),
        n_outlet=len(n["outlet"]["node_id"]),
        n_ud=len(n["user_demand"]["node_id"]),
        n_ud_link=sr["user_demand_inflow"][1] - sr["user_demand_inflow"][0],
        n_lr=len(n["linear_resistance"]["node_id"]),
        n_manning=len(n["manning_resistance"]["node_id"]),
        n_lb=len(n["level_boundary"]["node_id"]),
        n_fb=len(n["flow_boundary"]["node_id"]),
        n_cc=len(n["continuous_control"]["node_id"]),
        off_trc=sr["tabulated_rating_curve"][0],
        off_pump=sr["pump"][0],
        off_outlet=sr["outlet"][0],
        off_ud_in=sr["user_demand_inflow"][0],
        off_ud_out=sr["user_demand_outflow"][0],
        off_lr=sr["linear_resistance"][0],
        off_manning=sr["manning_resistance"][0],
        off_evap=sr["evaporation"][0],
        off_infil=sr["infiltration"][0],
        off_integral=sr["integral"][0],
    )
    f.floats["level_difference_threshold"] = float(p.level_difference_threshold)
    A = f.arrays
    A["state_node_id"] = as_i32([s.value for s in p.state_node_id])

    # ------------------------------------------------------------------ basins
    A["storage0"] = as_f64(basin.storage0)
    A["low_storage_threshold"] = as_f64(basin.low_storage_threshold)
    A["fixed_area"] = as_f64([a[-1] for a in basin.areas])  # basin_areas(basin, i)[end]
    A["prof_ptr"], A["prof_level"] = _ragged(basin.levels)
    _, A["prof_dSdh"] = _ragged(basin.dSdh)
    _, A["prof_cumS"] = _ragged(basin.cumS)
    _, A["prof_area"] = _ragged(basin.areas)
    # per-segment slopes, padded to the knot count so all tables share prof_ptr
    _, A["prof_dSdh_slope"] = _ragged([s + [0.0] for s in basin.dSdh_slope])
    _, A["prof_area_slope"] = _ragged([s + [0.0] for s in basin.area_slope])

    # ------------------------------------------------------------------ incidence (reduce_state!)
    def state_of_link(a: NodeID, b: NodeID) -> int | None:
        # get_state_index(state_ranges, link_to_state_idx, link) (util.jl:904-913)
        idx = p.link_to_state_idx.get((a, b))
        if idx is not None:
            return idx
        idx = _node_state(p, b, inflow=True)
        if idx is not None:
            return idx
        return _node_state(p, a, inflow=False)

    inc_state: list[list[int]] = []
    inc_sign: list[list[float]] = []
    for i, bid in enumerate(basin.node_id):
        st, sg = [], []
        for src in basin.inflow_ids[i]:
            s = state_of_link(src, bid)
            if s is None:
                s = _node_state(p, src, inflow=False)
            if s is None:
                continue
            st.append(s)
            sg.append(1.0)
        for dst in basin.outflow_ids[i]:
            s = state_of_link(bid, dst)
            if s is None:
                s = _node_state(p, dst, inflow=True)
            if s is None:
                continue
            st.append(s)
            sg.append(-1.0)
        inc_state.append(st)
        inc_sign.append(sg)
    A["inc_ptr"], A["inc_state"] = _ragged(inc_state, np.int32)
    _, A["inc_sign"] = _ragged(inc_sign, np.float64)

    # flow boundaries feeding basins (formulate_storages!, solve.jl:192-202)
    fb_basin = []
    for lm in n["flow_boundary"]["outflow_link"]:
        dst = lm.link[1]
        fb_basin.append(dst.idx - 1 if dst.type == "Basin" else -1)
    A["fb_basin"] = as_i32(fb_basin)
    fb_of_basin: list[list[int]] = [[] for _ in range(n_b)]
    for k, b in enumerate(fb_basin):
        if b >= 0:
            fb_of_basin[b].append(k)
    A["bfb_ptr"], A["bfb_idx"] = _ragged(fb_of_basin, np.int32)

    # ------------------------------------------------------------------ connector states
    conn_type = np.zeros(n_conn, np.int32)
    conn_idx = np.zeros(n_conn, np.int32)
    src_kind = np.full(n_conn, KIND_NONE, np.int32)
    src_idx = np.zeros(n_conn, np.int32)
    dst_kind = np.full(n_conn, KIND_NONE, np.int32)
    dst_idx = np.zeros(n_conn, np.int32)
    comp_ct = {
        "tabulated_rating_curve": CT_TRC, "pump": CT_PUMP, "outlet": CT_OUTLET,
        "user_demand_inflow": CT_UD_IN, "user_demand_outflow": CT_UD_OUT,
        "linear_resistance": CT_LR, "manning_resistance": CT_MANNING,
    }
    for comp, ct in comp_ct.items():
        a, b = sr[comp]
        for s in range(a, b):
            conn_type[s] = ct
            conn_idx[s] = s - a
            li, lo = p.state_inflow_link[s], p.state_outflow_link[s]
            if li is not None:
                src_kind[s], src_idx[s] = _endpoint(li.link[0])
            if lo is not None:
                dst_kind[s], dst_idx[s] = _endpoint(lo.link[1])
    A.update(conn_type=conn_type, conn_idx=conn_idx, src_kind=src_kind, src_idx=src_idx,
             dst_kind=dst_kind, dst_idx=dst_idx)

    # ------------------------------------------------------------------ static link parameters
    lr = n["linear_resistance"]
    A["lr_resistance"] = as_f64(lr["resistance"])
    A["lr_max_flow_rate"] = as_f64(lr["max_flow_rate"])
    mn = n["manning_resistance"]
    for k in ("length", "manning_n", "profile_width", "profile_slope", "upstream_bottom",
              "downstream_bottom"):
        A[f"mn_{k}"] = as_f64(mn[k])
    trc = n["tabulated_rating_curve"]
    tabs = list(trc["interpolations"]) + list(n["continuous_control"]["func"])
    f.ints["n_tab"] = len(tabs)
    f.ints["cc_tab_offset"] = len(trc["interpolations"])
    A["tab_ptr"], A["tab_t"] = _ragged([t.t for t in tabs])
    _, A["tab_u"] = _ragged([t.u for t in tabs])
    _, A["tab_du"] = _ragged([t.du for t in tabs])
    _, A["tab_c1"] = _ragged([t.c1 + [0.0] for t in tabs])
    _, A["tab_c2"] = _ragged([t.c2 + [0.0] for t in tabs])
    A["tab_slope_left"] = as_f64([t.slope_left for t in tabs])
    A["tab_slope_right"] = as_f64([t.slope_right for t in tabs])

    for key, pre in (("pump", "pump"), ("outlet", "outlet")):
        A[f"{pre}_control_type"] = as_i32(n[key]["control_type"])
    ud = n["user_demand"]
    A["ud_link_ptr"] = as_i32(ud["inflow_link_offsets"])
    A["ud_min_level"] = as_f64(ud["min_level"])
    A["ud_link_alloc"] = as_f64([x for l in ud["inflow_link_allocated"] for x in l])
    ud_of_link = []
    for i, links in enumerate(ud["inflow_links"]):
        ud_of_link.extend([i] * len(links))
    A["ud_of_link"] = as_i32(ud_of_link)

    # ------------------------------------------------------------------ PID / continuous control
    pid = n["pid_control"]
    A["pid_listen_basin"] = as_i32([l.idx - 1 for l in pid["listen_node_id"]])
    A["pid_target_is_outlet"] = as_i32([1 if t.type == "Outlet" else 0 for t in pid["target_node"]])
    A["pid_target_idx"] = as_i32([t.idx - 1 for t in pid["target_node"]])
    cc = n["continuous_control"]
    A["cc_target_is_outlet"] = as_i32([1 if t.type == "Outlet" else 0 for t in cc["target_node"]])
    A["cc_target_idx"] = as_i32([t.idx - 1 for t in cc["target_node"]])
    sv_kind, sv_idx, sv_w = [], [], []
    for cvars in cc["compound_variable"]:
        sv_kind.append([sv["kind"] for sv in cvars])
        sv_idx.append([sv["idx"] for sv in cvars])
        sv_w.append([sv["weight"] for sv in cvars])
    A["cc_sv_ptr"], A["cc_sv_kind"] = _ragged(sv_kind, np.int32)
    _, A["cc_sv_idx"] = _ragged(sv_idx, np.int32)
    _, A["cc_sv_weight"] = _ragged(sv_w, np.float64)

    _flatten_discrete_control(p, f)
    _build_patterns(p, f)
    return f


def _flatten_discrete_control(p: Params, f: Flat) -> None:
    """DiscreteControl (core/src/callback.jl:596-696) for per-member evaluation on the device.

    Conditions are numbered DC node by DC node in truth-state order (compound variable, then
    condition); a truth state is the bit pattern `sum(truth[j] << j)`; `dc_logic` maps it to the
    index of the control state in the node's own list of state names (-1: undefined).  Every
    controlled node gets one parameter row per control state of its DiscreteControl node."""
    n = p.nodes
    A = f.arrays
    dc = n["discrete_control"]
    n_dc = len(dc["node_id"])
    f.ints["n_dc"] = n_dc
    cond_ptr = [0]
    csv_ptr = [0]
    sv_kind: list[int] = []
    sv_idx: list[int] = []
    sv_w: list[float] = []
    logic_ptr = [0]
    logic: list[int] = []
    state_names: list[list[str]] = []
    for i in range(n_dc):
        n_cond = 0
        for cv in dc["compound_variables"][i]:
            for _ in cv["threshold_high"]:
                for sv in cv["subvariables"]:
                    sv_kind.append(sv["kind"])
                    sv_idx.append(sv["idx"])
                    sv_w.append(sv["weight"])
                csv_ptr.append(len(sv_kind))
                n_cond += 1
        cond_ptr.append(cond_ptr[-1] + n_cond)
        lm = dc["logic_mapping"][i]
        names = sorted(set(lm.values()))
        state_names.append(names)
        table = [-1] * (1 << n_cond)
        for truth, cs in lm.items():
            bits = sum(1 << j for j, b in enumerate(truth) if b)
            table[bits] = names.index(cs)
        logic.extend(table)
        logic_ptr.append(len(logic))
    f.ints["n_dc_cond"] = cond_ptr[-1]
    A["dc_cond_ptr"] = as_i32(cond_ptr)
    A["dc_csv_ptr"] = as_i32(csv_ptr)
    A["dc_sv_kind"] = as_i32(sv_kind)
    A["dc_sv_idx"] = as_i32(sv_idx)
    A["dc_sv_weight"] = as_f64(sv_w)
    A["dc_logic_ptr"] = as_i32(logic_ptr)
    A["dc_logic"] = as_i32(logic)
    A["dc_n_states"] = as_i32([len(x) for x in state_names])
    f.dc_state_names = state_names

    dc_of: dict[int, int] = {}
    for i, targets in enumerate(dc["controlled_nodes"]):
        for tnode in targets:
            dc_of[tnode.value] = i

    def rows(key: str, fields: list[str], base: dict[str, np.ndarray]):
        d = n[key]
        ctrl, ptr = [], [0]
        vals = {fl: [] for fl in fields}
        cm = d.get("control_mapping", {})
        for k, nid in enumerate(d["node_id"]):
            i = dc_of.get(nid.value, -1)
            ctrl.append(i)
            if i >= 0:
                for cs in state_names[i]:
                    upd = cm.get((nid.value, cs), {})
                    for fl in fields:
                        vals[fl].append(float(upd[fl]) if fl in upd else float(base[fl][k]))
            ptr.append(len(vals[fields[0]]))
        return as_i32(ctrl), as_i32(ptr), {fl: as_f64(v) for fl, v in vals.items()}

    for key, pre in (("pump", "pump"), ("outlet", "outlet")):
        d = n[key]
        base = dict(
            flow_rate=np.asarray(d["flow_rate"], dtype=np.float64),
            min_flow_rate=np.array([s(0.0) for s in d["min_flow_rate"]]),
            max_flow_rate=np.array([s(0.0) for s in d["max_flow_rate"]]),
            min_upstream_level=np.array([s(0.0) for s in d["min_upstream_level"]]),
            max_downstream_level=np.array([s(0.0) for s in d["max_downstream_level"]]),
        )
        ctrl, ptr, vals = rows(key, list(base), base)
        A[f"{pre}_dc"], A[f"{pre}_cs_ptr"] = ctrl, ptr
        for fl, v in vals.items():
            A[f"{pre}_cs_{fl}"] = v
    lr = n["linear_resistance"]
    ctrl, ptr, vals = rows("linear_resistance", ["resistance", "max_flow_rate"],
                           dict(resistance=lr["resistance"], max_flow_rate=lr["max_flow_rate"]))
    A["lr_dc"], A["lr_cs_ptr"] = ctrl, ptr
    A["lr_cs_resistance"], A["lr_cs_max_flow_rate"] = vals["resistance"], vals["max_flow_rate"]
    mn = n["manning_resistance"]
    mn_fields = ["length", "manning_n", "profile_width", "profile_slope"]
    ctrl, ptr, vals = rows("manning_resistance", mn_fields, {fl: mn[fl] for fl in mn_fields})
    A["mn_dc"], A["mn_cs_ptr"] = ctrl, ptr
    for fl in mn_fields:
        A[f"mn_cs_{fl}"] = vals[fl]
    pid = n["pid_control"]
    pid_fields = ["target", "proportional", "integral", "derivative"]
    ctrl, ptr, vals = rows("pid_control", pid_fields,
                           {fl: np.array([float(s(0.0)) for s in pid[fl]]) for fl in pid_fields})
    A["pid_dc"], A["pid_cs_ptr"] = ctrl, ptr
    for fl in pid_fields:
        A[f"pid_cs_{fl}"] = vals[fl]
    trc = n["tabulated_rating_curve"]
    base_tab = np.array([float(s(0.0)) for s in trc["current_interpolation_index"]])
    ctrl, ptr, vals = rows("tabulated_rating_curve", ["table"], dict(table=base_tab))
    A["trc_dc"], A["trc_cs_ptr"] = ctrl, ptr
    A["trc_cs_table"] = as_i32(vals["table"] - 1)  # 0-based table index


def _node_state(p: Params, nid: NodeID, inflow: bool) -> int | None:
    """get_state_index(state_ranges, id; inflow) (util.jl:878-895)."""
    comp = {
        "TabulatedRatingCurve": "tabulated_rating_curve", "Pump": "pump", "Outlet": "outlet",
        "LinearResistance": "linear_resistance", "ManningResistance": "manning_resistance",
        "UserDemand": "user_demand_inflow" if inflow else "user_demand_outflow",
    }.get(nid.type)
    if comp is None:
        return None
    a, b = p.state_ranges[comp]
    if comp == "user_demand_inflow":
        # one state per inflow link; the node-indexed lookup resolves to the idx-th entry
        k = a + nid.idx - 1
        return k if k < b else None
    return a + nid.idx - 1


def _build_patterns(p: Params, f: Flat) -> None:
    A = f.arrays
    n_b, n_pid = f.ints["n_basin"], f.ints["n_pid"]
    n_state, n_conn = f.ints["n_state"], f.ints["n_conn"]
    n_r = n_b + n_pid
    sr = p.state_ranges
    src_kind, src_idx = A["src_kind"], A["src_idx"]
    dst_kind, dst_idx = A["dst_kind"], A["dst_idx"]
    conn_type = A["conn_type"]
    nodes = p.nodes

    rows: list[list[int]] = [[] for _ in range(n_state)]

    def add(r: int, c: int) -> None:
        if c not in rows[r]:
            rows[r].append(c)

    # base connector rows
    for s in range(n_conn):
        if conn_type[s] == CT_UD_OUT:
            continue
        if src_kind[s] == KIND_BASIN:
            add(s, int(src_idx[s]))
        if dst_kind[s] == KIND_BASIN:
            add(s, int(dst_idx[s]))
    # user demand outflow depends on every inflow link's source
    ud = nodes["user_demand"]
    for i in range(len(ud["node_id"])):
        r = sr["user_demand_outflow"][0] + i
        for k in range(ud["inflow_link_offsets"][i], ud["inflow_link_offsets"][i + 1]):
            s = sr["user_demand_inflow"][0] + k
            if src_kind[s] == KIND_BASIN:
                add(r, int(src_idx[s]))
    for b in range(n_b):
        add(sr["evaporation"][0] + b, b)
        add(sr["infiltration"][0] + b, b)

    # continuous control: target row inherits the columns of what it listens to
    cc = nodes["continuous_control"]
    for i, tgt in enumerate(cc["target_node"]):
        comp = "outlet" if tgt.type == "Outlet" else "pump"
        r = sr[comp][0] + tgt.idx - 1
        for sv in cc["compound_variable"][i]:
            if sv["kind"] in (SV_BASIN_LEVEL, SV_BASIN_STORAGE):
                add(r, sv["idx"])
            elif sv["kind"] == SV_DU_STATE:
                for c in list(rows[sv["idx"]]):
                    add(r, c)

    # PID
    pid = nodes["pid_control"]
    inc_ptr, inc_state = A["inc_ptr"], A["inc_state"]
    pid_has_kd = []
    for i, tgt in enumerate(pid["target_node"]):
        comp = "outlet" if tgt.type == "Outlet" else "pump"
        r = sr[comp][0] + tgt.idx - 1
        L = pid["listen_node_id"][i].idx - 1
        add(sr["integral"][0] + i, L)
        add(r, L)
        add(r, n_b + i)
        kd_vals = list(pid["derivative"][i].u)
        for (nv, _cs), upd in pid["control_mapping"].items():
            if nv == pid["node_id"][i].value and "derivative" in upd:
                kd_vals.append(upd["derivative"])
        has_kd = any(v != 0.0 for v in kd_vals)
        pid_has_kd.append(1 if has_kd else 0)
        if has_kd:
            for k in range(inc_ptr[L], inc_ptr[L + 1]):
                s = int(inc_state[k])
                # flows of PID-controlled nodes are still zero when the PID term is formed
                if _is_pid_controlled(p, s):
                    continue
                for c in list(rows[s]):
                    add(r, c)
    A["pid_has_kd"] = as_i32(pid_has_kd)

    for r in range(n_state):
        rows[r].sort()
    jrow_ptr = [0]
    jcol: list[int] = []
    for r in range(n_state):
        jcol.extend(rows[r])
        jrow_ptr.append(len(jcol))
    A["jrow_ptr"] = as_i32(jrow_ptr)
    A["jcol"] = as_i32(jcol)
    f.ints["nnz_j"] = len(jcol)
    # CSC view (rows ascending inside a column), 1-based like the golden vectors
    pairs = sorted((c, r) for r in range(n_state) for c in rows[r])
    f.jac_rows_csc = as_i32([r + 1 for c, r in pairs])
    f.jac_cols_csc = as_i32([c + 1 for c, r in pairs])

    def jpos(r: int, c: int) -> int:
        return jrow_ptr[r] + rows[r].index(c)

    # positions of the direct (own endpoint) partials of each connector row
    jpos_src = np.full(n_conn, -1, np.int32)
    jpos_dst = np.full(n_conn, -1, np.int32)
    for s in range(n_conn):
        if conn_type[s] == CT_UD_OUT:
            continue
        if src_kind[s] == KIND_BASIN:
            jpos_src[s] = jpos(s, int(src_idx[s]))
        if dst_kind[s] == KIND_BASIN:
            jpos_dst[s] = jpos(s, int(dst_idx[s]))
    A["jpos_src"], A["jpos_dst"] = jpos_src, jpos_dst
    # UD outflow: for every inflow link, the slot of its source basin in the outflow row
    ud_out_pos = []
    for i in range(len(ud["node_id"])):
        r = sr["user_demand_outflow"][0] + i
        for k in range(ud["inflow_link_offsets"][i], ud["inflow_link_offsets"][i + 1]):
            s = sr["user_demand_inflow"][0] + k
            ud_out_pos.append(jpos(r, int(src_idx[s])) if src_kind[s] == KIND_BASIN else -1)
    A["ud_out_jpos"] = as_i32(ud_out_pos)

    # chain terms for controlled rows: (target position, source position, subvariable index)
    # continuous control
    cc_term_ptr = [0]
    cc_term_kind: list[int] = []  # 0: basin level, 1: basin storage, 2: copy of J entry
    cc_term_sv: list[int] = []
    cc_term_src: list[int] = []
    cc_term_dst: list[int] = []
    sv_base = 0
    for i, tgt in enumerate(cc["target_node"]):
        comp = "outlet" if tgt.type == "Outlet" else "pump"
        r = sr[comp][0] + tgt.idx - 1
        for j, sv in enumerate(cc["compound_variable"][i]):
            if sv["kind"] == SV_BASIN_LEVEL:
                cc_term_kind.append(0); cc_term_sv.append(sv_base + j)
                cc_term_src.append(sv["idx"]); cc_term_dst.append(jpos(r, sv["idx"]))
            elif sv["kind"] == SV_BASIN_STORAGE:
                cc_term_kind.append(1); cc_term_sv.append(sv_base + j)
                cc_term_src.append(sv["idx"]); cc_term_dst.append(jpos(r, sv["idx"]))
            elif sv["kind"] == SV_DU_STATE:
                s = sv["idx"]
                for c in rows[s]:
                    cc_term_kind.append(2); cc_term_sv.append(sv_base + j)
                    cc_term_src.append(jpos(s, c)); cc_term_dst.append(jpos(r, c))
        sv_base += len(cc["compound_variable"][i])
        cc_term_ptr.append(len(cc_term_kind))
    A["cc_term_ptr"] = as_i32(cc_term_ptr)
    A["cc_term_kind"] = as_i32(cc_term_kind)
    A["cc_term_sv"] = as_i32(cc_term_sv)
    A["cc_term_src"] = as_i32(cc_term_src)
    A["cc_term_dst"] = as_i32(cc_term_dst)
    cc_row_ptr = []
    for i, tgt in enumerate(cc["target_node"]):
        comp = "outlet" if tgt.type == "Outlet" else "pump"
        cc_row_ptr.append(sr[comp][0] + tgt.idx - 1)
    A["cc_target_state"] = as_i32(cc_row_ptr)

    # PID: slots and derivative-term lists
    pid_state, pid_jpos_listen, pid_jpos_integral, pid_int_jpos = [], [], [], []
    pid_term_ptr = [0]
    pid_term_state: list[int] = []   # du state entering dS_listened
    pid_term_sign: list[float] = []
    pid_jterm_ptr = [0]
    pid_jterm_src: list[int] = []    # J entry of an adjacent row
    pid_jterm_dst: list[int] = []    # slot in the target row
    pid_jterm_sign: list[float] = []
    for i, tgt in enumerate(pid["target_node"]):
        comp = "outlet" if tgt.type == "Outlet" else "pump"
        r = sr[comp][0] + tgt.idx - 1
        L = pid["listen_node_id"][i].idx - 1
        pid_state.append(r)
        pid_jpos_listen.append(jpos(r, L))
        pid_jpos_integral.append(jpos(r, n_b + i))
        pid_int_jpos.append(jpos(sr["integral"][0] + i, L))
        for k in range(inc_ptr[L], inc_ptr[L + 1]):
            s = int(inc_state[k])
            sign = float(A["inc_sign"][k])
            pid_term_state.append(s)
            pid_term_sign.append(sign)
            if A["pid_has_kd"][i] and not _is_pid_controlled(p, s):
                for c in rows[s]:
                    pid_jterm_src.append(jpos(s, c))
                    pid_jterm_dst.append(jpos(r, c))
                    pid_jterm_sign.append(sign)
        pid_term_ptr.append(len(pid_term_state))
        pid_jterm_ptr.append(len(pid_jterm_src))
    A["pid_target_state"] = as_i32(pid_state)
    A["pid_jpos_listen"] = as_i32(pid_jpos_listen)
    A["pid_jpos_integral"] = as_i32(pid_jpos_integral)
    A["pid_int_jpos"] = as_i32(pid_int_jpos)
    A["pid_term_ptr"] = as_i32(pid_term_ptr)
    A["pid_term_state"] = as_i32(pid_term_state)
    A["pid_term_sign"] = as_f64(pid_term_sign)
    A["pid_jterm_ptr"] = as_i32(pid_jterm_ptr)
    A["pid_jterm_src"] = as_i32(pid_jterm_src)
    A["pid_jterm_dst"] = as_i32(pid_jterm_dst)
    A["pid_jterm_sign"] = as_f64(pid_jterm_sign)

    # ------------------------------------------------------------------ W_inner = A*J_i - I/gamma
    # contributions of J_i[r, c] to J_inner rows, by kind of r (differentiation.jl:189-262)
    contrib: dict[tuple[int, int], list[tuple[int, float]]] = {}
    for r in range(n_state):
        targets: list[tuple[int, float]] = []
        if r < n_conn:
            ct = conn_type[r]
            do_inflow = ct != CT_UD_OUT
            if do_inflow and src_kind[r] == KIND_BASIN:
                targets.append((int(src_idx[r]), -1.0))
            if ct != CT_UD_IN and dst_kind[r] == KIND_BASIN:
                targets.append((int(dst_idx[r]), 1.0))
        elif r < sr["infiltration"][1]:
            b = (r - sr["evaporation"][0]) % n_b
            targets.append((b, -1.0))
        else:
            targets.append((n_b + (r - sr["integral"][0]), 1.0))
        for c in rows[r]:
            for (i, sign) in targets:
                contrib.setdefault((i, c), []).append((jpos(r, c), sign))
    for i in range(n_r):
        contrib.setdefault((i, i), [])
    keys = sorted(contrib)  # CSR order (row, col)
    wrow_ptr = [0] * (n_r + 1)
    wcol, wsrc_ptr, wsrc_pos, wsrc_sign, wdiag = [], [0], [], [], []
    for (i, c) in keys:
        wrow_ptr[i + 1] += 1
        if i == c:
            wdiag.append(len(wcol))
        wcol.append(c)
        # contributions arrive in ascending state row, as in calc_J_inner!'s column sweep
        for (pos, sign) in contrib[(i, c)]:
            wsrc_pos.append(pos)
            wsrc_sign.append(sign)
        wsrc_ptr.append(len(wsrc_pos))
    for i in range(n_r):
        wrow_ptr[i + 1] += wrow_ptr[i]
    A["wrow_ptr"] = as_i32(wrow_ptr)
    A["wcol"] = as_i32(wcol)
    A["wdiag"] = as_i32(wdiag)
    A["wsrc_ptr"] = as_i32(wsrc_ptr)
    A["wsrc_pos"] = as_i32(wsrc_pos)
    A["wsrc_sign"] = as_f64(wsrc_sign)
    f.ints["nnz_w"] = len(wcol)


def _is_pid_controlled(p: Params, s: int) -> bool:
    for comp in ("pump", "outlet"):
        a, b = p.state_ranges[comp]
        if a <= s < b:
            return p.nodes[comp]["control_type"][s - a] == CONTROL_PID
    return False
