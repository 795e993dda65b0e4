"""Host network compiler against the reference's own golden values (no GPU).

Everything here is integer layout: state order, incidence matrix A, Jacobian sparsity,
W_inner assembly lists, symbolic LU.  Citations are relative to /root/reference.
"""

from __future__ import annotations

import numpy as np
import pytest

from ribasim_b200.flatten import flatten
from ribasim_b200.network import build_params
from ribasim_b200.symbolic import (analyse, build_entries, numeric_entries, numeric_factor,
                                   numeric_factor_ldu, numeric_solve, numeric_solve_ldu)
from ribasim_b200.tables import ModelTables
from ribasim_b200.testmodels import MODELS


@pytest.fixture(scope="module")
def basic():
    p = build_params(MODELS["basic"]())
    return p, flatten(p)


@pytest.fixture(scope="module")
def pid():
    p = build_params(MODELS["pid_control"]())
    return p, flatten(p)


def dense_A(flat) -> np.ndarray:
    nb, n, nr = flat["n_basin"], flat["n_state"], flat["n_reduced"]
    A = np.zeros((nr, n))
    for b in range(nb):
        for k in range(flat["inc_ptr"][b], flat["inc_ptr"][b + 1]):
            A[b, flat["inc_state"][k]] += flat["inc_sign"][k]
        A[b, flat["off_evap"] + b] -= 1.0
        A[b, flat["off_infil"] + b] -= 1.0
    for i in range(flat["n_pid"]):
        A[nb + i, flat["off_integral"] + i] = 1.0
    return A


def test_basic_state_node_ids(basic):
    # core/test/run_models_test.jl:217
    p, flat = basic
    assert [s.value for s in p.state_node_id] == [4, 5, 8, 7, 10, 12, 2, 1, 3, 6, 9, 1, 3, 6, 9]
    assert flat["n_basin"] == 4 and flat["n_conn"] == 7 and flat["n_state"] == 15


def test_trivial_state_axes():
    # core/test/run_models_test.jl:17-20: one TRC state + evaporation + infiltration
    p = build_params(MODELS["trivial"]())
    sr = p.state_ranges
    assert sr["tabulated_rating_curve"] == (0, 1)
    assert sr["evaporation"] == (1, 2) and sr["infiltration"] == (2, 3)
    assert p.n_states == 3


def test_basic_jacobian_sparsity(basic):
    # core/test/utils_test.jl:266-271
    _, flat = basic
    rows = [7, 8, 12, 1, 2, 3, 6, 7, 9, 13, 2, 4, 10, 14, 3, 4, 5, 11, 15]
    cols = [1, 1, 1, 2, 2, 2, 2, 2, 2, 2, 3, 3, 3, 3, 4, 4, 4, 4, 4]
    assert list(flat.jac_rows_csc) == rows
    assert list(flat.jac_cols_csc) == cols
    assert flat["nnz_j"] == 19


def test_pid_jacobian_sparsity(pid):
    # core/test/utils_test.jl:287-292
    _, flat = pid
    assert list(flat.jac_rows_csc) == [1, 2, 3, 4, 1]
    assert list(flat.jac_cols_csc) == [1, 1, 1, 1, 2]


def test_basic_incidence_matrix(basic):
    # core/test/utils_test.jl:423-428
    _, flat = basic
    rows = [2, 2, 3, 2, 4, 3, 4, 4, 2, 1, 2, 1, 2, 3, 4, 1, 2, 3, 4]
    cols = [1, 2, 2, 3, 3, 4, 4, 5, 6, 7, 7, 8, 9, 10, 11, 12, 13, 14, 15]
    vals = [-1.0, -1.0, 1.0, -1.0, 1.0, -1.0, 1.0, -1.0, 1.0, -1.0, 1.0, -1.0, -1.0, -1.0, -1.0,
            -1.0, -1.0, -1.0, -1.0]
    A_expected = np.zeros((4, 15))
    for r, c, v in zip(rows, cols, vals):
        A_expected[r - 1, c - 1] = v
    assert np.array_equal(dense_A(flat), A_expected)


def test_pid_incidence_matrix(pid):
    # core/test/utils_test.jl:445-450
    _, flat = pid
    A_expected = np.zeros((2, 4))
    for r, c, v in zip([1, 1, 1, 2], [1, 2, 3, 4], [-1.0, -1.0, -1.0, 1.0]):
        A_expected[r - 1, c - 1] = v
    assert np.array_equal(dense_A(flat), A_expected)


def _w_from_lists(flat, jvals_csr: np.ndarray, inv_gamma: float) -> np.ndarray:
    """What k_assemble_w computes, on the host."""
    nr = flat["n_reduced"]
    W = np.zeros((nr, nr))
    rp, col = flat["wrow_ptr"], flat["wcol"]
    diag = set(int(x) for x in flat["wdiag"])
    e = 0
    for i in range(nr):
        for e in range(rp[i], rp[i + 1]):
            acc = 0.0
            for k in range(flat["wsrc_ptr"][e], flat["wsrc_ptr"][e + 1]):
                acc += flat["wsrc_sign"][k] * jvals_csr[flat["wsrc_pos"][k]]
            if e in diag:
                acc -= inv_gamma
            W[i, col[e]] = acc
    return W


def _csr_from_csc_vals(flat, vals_csc) -> tuple[np.ndarray, np.ndarray]:
    n, nr = flat["n_state"], flat["n_reduced"]
    J = np.zeros((n, nr))
    for r, c, v in zip(flat.jac_rows_csc, flat.jac_cols_csc, vals_csc):
        J[r - 1, c - 1] = v
    rp, col = flat["jrow_ptr"], flat["jcol"]
    csr = np.array([J[r, col[e]] for r in range(n) for e in range(rp[r], rp[r + 1])])
    return J, csr


def test_basic_J_inner_is_A_times_J(basic):
    # core/test/utils_test.jl:430-435 with the reference's arbitrary nzval
    _, flat = basic
    nz = [0.020047016741082002, 0.8755256160248737, 0.36909649531559285, 0.7632298275012108,
          0.9240314657308235, 0.49544793385910524, 0.10528087709131306, 0.020608175445295474,
          0.9691738934605421, 0.4218954216679456, 0.5058554068921941, 0.2896077753195684,
          0.8694315735708924, 0.8458965765906646, 0.7966585871607135, 0.2581915440964345,
          0.6505806124461845, 0.8411882038236067, 0.8067685192045705]
    J, csr = _csr_from_csc_vals(flat, nz)
    W = _w_from_lists(flat, csr, 0.0)
    assert np.allclose(W, dense_A(flat) @ J, rtol=1e-14, atol=0)
    gamma = 3600.0
    W = _w_from_lists(flat, csr, 1.0 / gamma)
    assert np.allclose(W, dense_A(flat) @ J - np.eye(4) / gamma, rtol=1e-14, atol=0)


def test_pid_J_inner_is_A_times_J(pid):
    # core/test/utils_test.jl:452-459
    _, flat = pid
    nz = [0.449381314574683, 0.4542317082538514, 0.599934409205972, 0.30717602583409154,
          0.9795352040440034]
    J, csr = _csr_from_csc_vals(flat, nz)
    assert np.allclose(_w_from_lists(flat, csr, 0.0), dense_A(flat) @ J, rtol=1e-14, atol=0)


@pytest.mark.parametrize("name", sorted(MODELS))
def test_patterns_are_consistent(name):
    """W pattern = pattern(A * J_pattern) + diagonal, for every re-expressed test model."""
    p = build_params(MODELS[name]())
    flat = flatten(p)
    n, nr = flat["n_state"], flat["n_reduced"]
    Jp = np.zeros((n, nr))
    rp, col = flat["jrow_ptr"], flat["jcol"]
    for r in range(n):
        cs = col[rp[r]:rp[r + 1]]
        assert list(cs) == sorted(set(int(c) for c in cs)), "CSR columns sorted and unique"
        Jp[r, cs] = 1.0
    Wp = (np.abs(dense_A(flat)) @ Jp > 0) | np.eye(nr, dtype=bool)
    got = np.zeros((nr, nr), dtype=bool)
    wrp, wcol = flat["wrow_ptr"], flat["wcol"]
    for i in range(nr):
        got[i, wcol[wrp[i]:wrp[i + 1]]] = True
    assert np.array_equal(got, Wp)


@pytest.mark.parametrize("name,kw", [("basic", {}), ("backwater", {}), ("pid_control", {}),
                                     ("outlet_continuous_control", {}), ("user_demand", {}),
                                     ("hws_like", dict(n_basins=150, n_days=1)),
                                     ("random_tree", dict(n_basins=400, n_hours=2))])
def test_symbolic_lu_solves(name, kw):
    from ribasim_b200.testmodels import hws_like_model, random_tree_model
    make = dict(MODELS, hws_like=hws_like_model, random_tree=random_tree_model)[name]
    p = build_params(make(**kw))
    flat = flatten(p)
    nr = flat["n_reduced"]
    sym = analyse(nr, flat["wrow_ptr"], flat["wcol"])
    rng = np.random.default_rng(3)
    nnz = flat["nnz_w"]
    wv = rng.normal(size=nnz)
    wv[flat["wdiag"]] -= 8.0  # diagonally dominant like W = J - I/gamma
    W = np.zeros((nr, nr))
    wrp, wcol = flat["wrow_ptr"], flat["wcol"]
    for i in range(nr):
        for e in range(wrp[i], wrp[i + 1]):
            W[i, wcol[e]] = wv[e]
    b = rng.normal(size=nr)
    lu = numeric_factor(sym, wv)
    x = numeric_solve(sym, lu, b)
    assert np.allclose(W @ x, b, rtol=1e-10, atol=1e-10)
    # LDU variant (one level per elimination round, reciprocal pivots)
    lud = numeric_factor_ldu(sym, wv)
    xd = numeric_solve_ldu(sym, lud, b)
    assert np.allclose(W @ xd, b, rtol=1e-10, atol=1e-10)
    assert len(sym.ldu_level_ptr) - 1 == sym.n_rounds
    # entry schedule of the tile kernels (emulation with the kernel's summation order)
    ent = build_entries(sym)
    xe = numeric_entries(sym, ent, wv, b)
    assert np.allclose(W @ xe, b, rtol=1e-10, atol=1e-10)
    assert ent.lvl_ptr[-1] == ent.ent_e.size and len(ent.team_log) == ent.n_elim + ent.n_bwd
    # race freedom of a level on the device: a slot is written by one entry only and no entry
    # reads a slot that the level writes (levels are separated by block barriers)
    # the compact form (fill-ins take over the slots of dead entries, full-Newton path): fewer
    # slots, the same bits — the emulation starts from NaN in every slot that is not loaded, so a
    # read of stale content would show
    entc = build_entries(sym, compact=True)
    assert entc.n_slots <= ent.n_slots and entc.y0 + nr == entc.n_slots
    assert np.array_equal(numeric_entries(sym, entc, wv, b), xe)
    assert len(set(entc.load_s.tolist())) == entc.load_s.size
    for ent in (ent, entc):
        for lv in range(ent.n_elim + ent.n_bwd):
            written, read = [], set()
            for k in range(ent.lvl_ptr[lv], ent.lvl_ptr[lv + 1]):
                written.append(int(ent.ent_e[k]))
                cnt, q0 = int(ent.ent_cf[k]) & 0xFF, int(ent.ent_q0[k])
                for q in range(q0, q0 + cnt):
                    read.update((int(ent.prod_l[q]), int(ent.prod_p[q]), int(ent.prod_u[q])))
                if int(ent.ent_cf[k]) & 0x200 and cnt == 0:
                    read.add(q0)  # the reciprocal pivot of a row without U entries
            assert len(written) == len(set(written)), lv
            assert not (read & set(written)), lv


def test_symbolic_lu_chain_is_logarithmic():
    """Rake-and-compress ordering: a 51-basin chain (tridiagonal W) factorises in O(log n)
    rounds with O(n) fill (cyclic reduction) instead of 51 sequential pivots."""
    p = build_params(MODELS["backwater"]())
    flat = flatten(p)
    sym = analyse(flat["n_reduced"], flat["wrow_ptr"], flat["wcol"])
    assert sym.n_rounds <= 8
    assert flat["nnz_w"] <= sym.n_lu <= 2 * flat["nnz_w"]
    assert len(sym.lu_level_ptr) - 1 <= 4 * sym.n_rounds


def test_tables_roundtrip(tmp_path):
    """TOML + GeoPackage(SQLite) writer/reader: same tables, same compiled network."""
    m = MODELS["basic"]()
    toml = m.write(tmp_path / "basic")
    m2 = ModelTables.read(toml)
    f1 = flatten(build_params(m))
    f2 = flatten(build_params(m2))
    assert f1.ints == f2.ints
    for k, v in f1.arrays.items():
        assert np.array_equal(v, f2.arrays[k], equal_nan=True), k


@pytest.mark.parametrize("char_ids", [False, True])
def test_netcdf_forcing_table_equals_database_table(tmp_path, char_ids):
    """`Basin / time` stored as a dense (time, node) NetCDF file named in the TOML (`[basin] time =
    "basin_time.nc"`, core/src/read.jl:1889-2098) compiles to the same network and forcing series
    as the same rows in the GeoPackage; ids as integers or as the Delft-FEWS char matrix."""
    from ribasim_b200.tables import load_netcdf, write_netcdf
    m = MODELS["basic_transient"]()
    ref = build_params(m)
    m.external_tables["Basin / time"] = "basin_time.nc"
    toml = m.write(tmp_path / "nc")
    if char_ids:
        write_netcdf(tmp_path / "nc" / m.input_dir / "basin_time.nc", "Basin / time",
                     m.table("Basin / time"), m.starttime, char_ids=True)
    assert "basin_time.nc" in toml.read_text()
    m2 = ModelTables.read(toml)
    assert m2.external_tables == {"Basin / time": "basin_time.nc"}
    rows = load_netcdf(tmp_path / "nc" / m.input_dir / "basin_time.nc", "Basin / time")
    assert len(rows) == len(m.table("Basin / time")) == 4 * 367
    assert rows[0]["node_id"] == 1 and rows[1]["node_id"] == 3  # node fastest, then time
    p2 = build_params(m2)
    f1, f2 = flatten(ref), flatten(p2)
    assert f1.ints == f2.ints
    for k, v in f1.arrays.items():
        assert np.array_equal(v, f2.arrays[k], equal_nan=True), k
    for name in ("precipitation", "potential_evaporation", "drainage", "infiltration"):
        for a, b in zip(ref.basin.forcing[name], p2.basin.forcing[name]):
            assert np.array_equal(a.t, b.t) and np.array_equal(a.u, b.u, equal_nan=True)
    assert ref.tstops == p2.tstops


def test_netcdf_drops_all_missing_rows(tmp_path):
    """Rows of the dense grid whose data variables are all missing do not exist in the table."""
    from datetime import datetime
    from ribasim_b200.tables import load_netcdf, write_netcdf
    t0 = datetime(2020, 1, 1)
    rows = [dict(node_id=7, time="2020-01-01 00:00:00", flow_rate=1.0),
            dict(node_id=9, time="2020-01-02 00:00:00", flow_rate=2.5)]
    write_netcdf(tmp_path / "fb.nc", "FlowBoundary / time", rows, t0)
    got = load_netcdf(tmp_path / "fb.nc", "FlowBoundary / time")
    assert [(r["node_id"], r["time"], r["flow_rate"]) for r in got] == \
        [(7, datetime(2020, 1, 1), 1.0), (9, datetime(2020, 1, 2), 2.5)]


def test_pump_discrete_control_compiles():
    p = build_params(MODELS["pump_discrete_control"]())
    flat = flatten(p)
    assert flat["n_pump"] >= 1


def test_junction_elimination_link_counts():
    # core/test/run_models_test.jl:672-701
    p = build_params(MODELS["junction_combined"]())
    assert len(p.internal_flow_links) == 8 and len(p.external_flow_links) == 10
    assert all(v.type != "Junction" for v in p.graph.code_of)
    p = build_params(MODELS["junction_chained"]())
    assert len(p.internal_flow_links) == 4 and len(p.external_flow_links) == 6
    assert len(p.graph.flow_link_pairs) == 8  # nnz(flow_link_map), run_models_test.jl:699-701
    assert all(v.type != "Junction" for v in p.graph.code_of)
    # both LinearResistance nodes flow out of Basin #1 through the eliminated junction chain
    lr = p.nodes["linear_resistance"]
    assert [l.link[0].value for l in lr["inflow_link"]] == [1, 1]
    assert [l.link[1].value for l in lr["outflow_link"]] == [6, 7]


def test_cyclic_tstops_goldens():
    # core/test/time_test.jl:75-88 (`get_timeseries_tstops`)
    from ribasim_b200.interp import TimeSeries
    assert TimeSeries([0, 0, 0], [0.5, 1.0, 1.5], kind="linear").tstops(5.0) == [0.5, 1.0, 1.5]
    assert np.array_equal(TimeSeries([0, 0, 0], [0.5, 1.0, 1.5], kind="linear", periodic=True).tstops(5.0),
                          np.arange(0, 5.01, 0.5))
    assert np.allclose(TimeSeries([0, 0], [0.3, 0.5], kind="constant", periodic=True).tstops(1.0),
                       np.arange(0.1, 0.91, 0.2), rtol=1e-12)


def test_cyclic_time_model():
    # core/test/time_test.jl:90-112: periodic extrapolation of the cyclic series; the precipitation
    # series contributes 404 tstops over the 101 years of the run
    from ribasim_b200.testmodels import cyclic_time_model
    p = build_params(cyclic_time_model())
    precip = p.basin.forcing["precipitation"][0]
    assert precip.periodic and p.nodes["level_boundary"]["level"][0].periodic
    assert p.nodes["flow_boundary"]["flow_rate"][0].periodic
    assert p.nodes["flow_boundary"]["flow_rate"][0].kind == "linear"
    assert len(precip.tstops(p.tables.t_end)) == 404


def test_validation_messages_of_the_reference():
    """core/test/validation_test.jl: neighbour counts (:46-99), Pump / Outlet flow rate sign
    (:196-236), link types (:238-262), TabulatedRatingCurve and Outlet upstream levels (:343-399),
    subgrid tables (:264-299) — same conditions, same messages."""
    from ribasim_b200.network import Graph, LinkMeta, NodeID, n_neighbor_errors
    from ribasim_b200.subgrid import Subgrid
    from ribasim_b200.tables import ModelTables

    # neighbour counts on the reference's hand-made graph
    g = Graph()
    ids = {k: NodeID(t, k, 1) for k, t in ((1, "Pump"), (2, "Basin"), (3, "Basin"), (4, "Basin"), (6, "Pump"))}
    for nid in ids.values():
        g.add_vertex(nid)
    for a, b in ((2, 1), (3, 1), (6, 2)):
        g.add_edge(ids[a], ids[b], LinkMeta(0, "flow", (ids[a], ids[b])))
    errs = [e for e in n_neighbor_errors(g) if e.startswith("Pump")]
    assert errs == ["Pump #1 can have at most 1 flow inneighbor(s) (got 2).",
                    "Pump #1 must have at least 1 flow outneighbor(s) (got 0).",
                    "Pump #6 must have at least 1 flow inneighbor(s) (got 0)."]

    def outlet_model(**static):
        m = MODELS["outlet"]()
        m.tables["Outlet / static"] = [dict(node_id=2, **static)]
        return m

    with pytest.raises(ValueError, match=r"Negative flow rate\(s\) for Outlet #2 found\."):
        build_params(outlet_model(flow_rate=-1.0))
    m = MODELS["flow_condition"]()
    m.tables["Pump / static"][1]["flow_rate"] = -1.0
    with pytest.raises(ValueError, match=r"Negative flow rate\(s\) found\."):
        build_params(m)

    m = MODELS["basic"]()
    m.link[0]["link_type"] = "foo"
    with pytest.raises(ValueError, match=r"Invalid link type 'foo' for link #1 from node #1 to node #2\."):
        build_params(m)

    m = MODELS["tabulated_rating_curve_control"]()
    for r in m.tables["TabulatedRatingCurve / static"]:
        if r["level"] == 0.0:
            r["level"] = -2.0
    with pytest.raises(ValueError, match="Lowest level of TabulatedRatingCurve #2 is lower than bottom of upstream Basin #1"):
        build_params(m)

    m = MODELS["level_boundary_condition"]()
    for r in m.tables["Outlet / static"]:
        r["min_upstream_level"] = -2.0
    with pytest.raises(ValueError, match="Minimum upstream level of Outlet #4 is lower than bottom of upstream Basin #3"):
        build_params(m)

    def subgrid_model(node_id, basin_level, subgrid_level):
        t = ModelTables(starttime=m.starttime, endtime=m.endtime)
        t.tables["Basin / subgrid"] = [dict(subgrid_id=1, node_id=node_id, basin_level=b, subgrid_level=s_)
                                       for b, s_ in zip(basin_level, subgrid_level)]
        return t

    with pytest.raises(ValueError, match="The node_id of the Basin / subgrid does not exist"):
        Subgrid(subgrid_model(10, [-1.0, 0.0], [-1.0, 0.0]), [9])
    with pytest.raises(ValueError, match="Basin / subgrid subgrid_id 1 has repeated basin levels, this cannot be interpolated."):
        Subgrid(subgrid_model(9, [-1.0, 0.0, 0.0], [-1.0, 0.0, 1.0]), [9])
    with pytest.raises(ValueError, match="Basin / subgrid subgrid_id 1 has repeated element levels, this cannot be interpolated."):
        Subgrid(subgrid_model(9, [-1.0, 0.0, 1.0], [-1.0, 0.0, 0.0]), [9])


def test_invalid_discrete_control_messages():
    """core/test/validation_test.jl:157-194 on the reference's `invalid_discrete_control` model
    (python/ribasim_testmodels/ribasim_testmodels/invalid.py:56-126): the five messages, in order."""
    from datetime import datetime

    from ribasim_b200.tables import ModelTables
    m = ModelTables(starttime=datetime(2020, 1, 1), endtime=datetime(2020, 12, 1))
    shared = dict(profile=dict(area=[0.01, 1.0], level=[0.0, 1.0]), state=dict(level=1.4112729908597084))
    m.add_node("Basin", 1, **shared)
    m.add_node("Pump", 2, static=dict(control_state="bar", flow_rate=0.5 / 3600))
    m.add_node("Basin", 3, **shared)
    m.add_node("FlowBoundary", 4, time=dict(time=["2020-01-01 00:00:00", "2020-11-01 00:00:00"],
                                            flow_rate=[1.0, 2.0]))
    m.add_node("DiscreteControl", 5,
               variable=dict(listen_node_id=[1, 4, 4], variable=["level", "flow_rate", "flow_rate"],
                             look_ahead=[100.0, 40 * 24 * 60 * 60.0, -10.0], compound_variable_id=[1, 2, 3]),
               condition=dict(threshold_high=[0.5, 1.5, 1.5], compound_variable_id=[1, 2, 3], condition_id=1),
               logic=dict(truth_state=["FFFF"], control_state=["foo"]))
    for a, b in [(1, 2), (2, 3), (4, 3), (5, 2)]:
        m.add_link(a, b)
    with pytest.raises(ValueError) as err:
        build_params(m)
    msg = str(err.value)
    expected = [
        'DiscreteControl #5 has 3 condition(s), which is inconsistent with these truth state(s): ["FFFF"].',
        'These control states from DiscreteControl #5 are not defined for controlled Pump #2: ["foo"].',
        "Look ahead supplied for non-timeseries listen variable 'level' from listen node Basin #1.",
        "Look ahead for listen variable 'flow_rate' from listen node FlowBoundary #4 goes past timeseries end during simulation.",
        "Negative look ahead supplied for listen variable 'flow_rate' from listen node FlowBoundary #4.",
    ]
    pos = [msg.find(e) for e in expected]
    assert all(p >= 0 for p in pos), (msg, pos)
    assert pos == sorted(pos)


def test_pid_connectivity_messages():
    # core/test/validation_test.jl:101-155
    m = MODELS["pid_control"]()
    m.tables["PidControl / time"] = [dict(r, listen_node_id=4) for r in m.tables["PidControl / time"]]
    with pytest.raises(ValueError, match="Listen node LevelBoundary #4 of PidControl #5 is not a Basin"):
        build_params(m)
    m = MODELS["pid_control"]()
    m.add_node("Basin", 9, profile=dict(area=[1000.0, 1000.0], level=[0.0, 1.0]), state=dict(level=1.0))
    m.tables["PidControl / time"] = [dict(r, listen_node_id=9) for r in m.tables["PidControl / time"]]
    with pytest.raises(ValueError, match=r"PID listened Basin #9 is not on either side of controlled Pump #3\."):
        build_params(m)


def test_time_helpers_goldens():
    # core/test/io_test.jl:41-51 (`datetime_since`, `seconds_since`) and :436-468 (`parse_time_strings!`)
    from datetime import datetime, timedelta

    from ribasim_b200.tables import datetime_since, parse_time, seconds_since
    t0 = datetime(2020, 1, 1)
    assert datetime_since(0.0, t0) == t0
    assert datetime_since(1.0, t0) == t0 + timedelta(seconds=1)
    assert datetime_since(np.pi, t0) == datetime(2020, 1, 1, 0, 0, 3, 142000)
    assert seconds_since(t0, t0) == 0.0
    assert seconds_since(t0 + timedelta(seconds=1), t0) == 1.0
    assert seconds_since(datetime(2020, 1, 1, 0, 0, 3, 142000), t0) == pytest.approx(3.142)
    assert parse_time("2020-06-01 12:00:00.000") == datetime(2020, 6, 1, 12)
    assert parse_time("2020-06-01 12:00:00.123456789") == datetime(2020, 6, 1, 12, 0, 0, 123000)
    assert parse_time("2020-06-01T12:00:00") == datetime(2020, 6, 1, 12)
