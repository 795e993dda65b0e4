"""The reference's result tables and files from a device run (ribasim_b200/results.py after
core/src/callback.jl:390-558 `save_flow` / `check_water_balance_error!` and core/src/write.jl)."""

from __future__ import annotations

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _model(name, dt=3600.0, **kw):
    from ribasim_b200.model import Model
    from ribasim_b200.network import build_params
    from ribasim_b200.testmodels import MODELS

    return Model(build_params(MODELS[name]()), dt=dt, **kw)


def test_basic_results_close_the_water_balance(tmp_path):
    """core/test/run_models_test.jl:224-276 (basic model results): one row per Basin and saved
    period, storages / levels at the start of the period, mean rates over it; the balance error
    of every Basin and period is within the reference's tolerances (it raises otherwise)."""
    from scipy.io import netcdf_file

    B = 3
    model = _model("basic", n_members=B)
    try:
        rng = np.random.default_rng(0)
        model.set_forcing("precipitation",
                          model.p.basin.vertical_flux["precipitation"][:, None] * rng.uniform(0.5, 2.0, (4, B)))
        S0 = model.storage()
        res = model.run(saveat=86400.0, results_dir=tmp_path)
        bd, fd = res.basin_data(), res.flow_data()
        n_t = 366
        assert bd["time"].shape == (n_t,) and bd["time"][1] == 86400.0
        assert bd["storage"].shape == (n_t, 4, B) and np.array_equal(bd["storage"][0], S0)
        # the storages of the next period's start are storage + storage_rate * saveat
        assert np.allclose(bd["storage"][1:], bd["storage"][:-1] + 86400.0 * bd["storage_rate"][:-1],
                           rtol=1e-12, atol=1e-9)
        # water balance of every basin, period and member (run_models_test.jl:259-271)
        assert np.all(np.abs(bd["balance_error"]) < 1e-3 * np.maximum(1.0, np.abs(bd["storage_rate"])))
        assert np.all(bd["inflow_rate"] >= 0) and np.all(bd["outflow_rate"] >= 0)
        assert np.all(bd["precipitation"] > 0) and np.all(bd["evaporation"] >= 0)
        # members differ through their precipitation only
        assert not np.array_equal(bd["precipitation"][..., 0], bd["precipitation"][..., 1])
        # flow table: the reference's external links (17 flow links in the basic model)
        assert fd["flow_rate"].shape == (n_t, len(model.p.external_flow_links), B)
        assert list(fd["link_id"]) == [l.id for l in model.p.external_flow_links]
        # a Basin's horizontal balance can be rebuilt from the link flows
        k = 100
        for b, nid in enumerate(bd["node_id"]):
            q_in = fd["flow_rate"][k][fd["to_node_id"] == nid].sum(axis=0)
            q_out = fd["flow_rate"][k][fd["from_node_id"] == nid].sum(axis=0)
            net = bd["inflow_rate"][k, b] - bd["outflow_rate"][k, b]
            assert np.allclose(q_in - q_out, net, rtol=1e-9, atol=1e-12)
        # files: names, dimensions and variables of write.jl
        with netcdf_file(str(tmp_path / "basin.nc"), "r", mmap=False) as ds:
            assert ds.variables["storage"].dimensions == ("realization", "time", "node_id")
            assert ds.variables["storage"].shape == (B, n_t, 4)
            assert np.array_equal(ds.variables["node_id"][:], bd["node_id"])
            assert np.array_equal(ds.variables["level"][:][1], bd["level"][..., 1])
            assert ds.variables["level"].units == b"m"
            for name in ("inflow_rate", "outflow_rate", "storage_rate", "precipitation", "surface_runoff",
                         "evaporation", "drainage", "infiltration", "balance_error", "relative_error"):
                assert name in ds.variables
        with netcdf_file(str(tmp_path / "flow.nc"), "r", mmap=False) as ds:
            assert ds.variables["flow_rate"].dimensions == ("realization", "time", "link_id")
            assert np.array_equal(ds.variables["from_node_id"][:], fd["from_node_id"])
        with netcdf_file(str(tmp_path / "basin_state.nc"), "r", mmap=False) as ds:
            assert np.array_equal(ds.variables["level"][:], model.level().T)
    finally:
        model.finalize()


def test_junction_flows_are_summed_over_external_links():
    """core/test/run_models_test.jl:665-704: with Junctions the external links carry the sum of
    the internal links routed over them; flow is conserved across every Junction."""
    model = _model("junction_combined", n_members=1)
    try:
        res = model.run(saveat=5 * 86400.0)
        fd = res.flow_data()
        q = fd["flow_rate"][-1, :, 0]
        junctions = [n.value for n in model.p.node_ids_all if n.type == "Junction"]
        assert junctions
        for j in junctions:
            assert q[fd["to_node_id"] == j].sum() == pytest.approx(q[fd["from_node_id"] == j].sum(), rel=1e-12)
        assert np.any(q != 0.0)
    finally:
        model.finalize()


def test_water_balance_error_is_raised():
    from ribasim_b200.results import Results, WaterBalanceError

    model = _model("basic", n_members=1)
    try:
        res = Results(model)
        model.update_until(86400.0)
        res._prev["storage"] = res._prev["storage"] + 1.0e4  # a storage the flows cannot explain
        with pytest.raises(WaterBalanceError, match="Too large water balance error"):
            res.save()
    finally:
        model.finalize()


def test_mean_flow_for_every_saveat():
    """core/test/run_models_test.jl:525-591 (`mean_flow`): the link from the constant FlowBoundary
    carries a mean flow of exactly 1 for every choice of saveat; one row per (started) period.  The
    linearly interpolated boundary is integrated exactly (solve.jl:86-99)."""
    dt = 24 * 24 * 60.0  # the solver step of the reference test
    t_end = 3.16224e7  # 366 days
    for saveat, n_rows in ((86400.0, 366), (float(round(10000 * np.pi)), int(np.ceil(t_end / round(10000 * np.pi)))),
                           (float("inf"), 1)):
        model = _model("flow_boundary_time", dt=dt, n_members=1)
        try:
            assert model.t_end == t_end
            res = model.run(saveat=saveat)
            fd = res.flow_data()
            flow = fd["flow_rate"][:, (fd["from_node_id"] == 3) & (fd["to_node_id"] == 2), 0].ravel()
            assert len(flow) == n_rows and len(res.times) == n_rows + 1
            assert np.allclose(flow, 1.0, rtol=1e-12, atol=0.0)
            # the time-varying boundary: mean over the period of the piecewise-linear series
            series = model.p.nodes["flow_boundary"]["flow_rate"][0]
            q1 = fd["flow_rate"][:, (fd["from_node_id"] == 1) & (fd["to_node_id"] == 2), 0].ravel()
            ts = np.array(res.times)
            grid = np.unique(np.concatenate([ts, np.asarray(series.t, dtype=float)]))
            grid = grid[(grid >= 0) & (grid <= t_end)]
            vals = np.array([series(float(x)) for x in grid])
            cum = np.concatenate([[0.0], np.cumsum(0.5 * (vals[1:] + vals[:-1]) * np.diff(grid))])
            exact = np.diff(np.interp(ts, grid, cum)) / np.diff(ts)
            assert np.allclose(q1, exact, rtol=1e-9)
            bd = res.basin_data()
            assert np.all(np.abs(bd["balance_error"]) < 1e-6)
        finally:
            model.finalize()


def test_two_basin_subgrid_levels(tmp_path):
    """core/test/run_models_test.jl:638-663 (`two_basin`): daily subgrid levels including the start,
    0.01 m at t0; the time-dependent h(h) relation of subgrid 2 rises by a metre after a month;
    apart from that shift the relations are 1:1."""
    from ribasim_b200.bmi import RibasimBmi
    from ribasim_b200.testmodels import MODELS

    model = _model("two_basin", n_members=2)
    try:
        res = model.run(saveat=86400.0, results_dir=tmp_path)
        sd = res.subgrid_level_data()
        ntime = 367
        assert sd["subgrid_level"].shape == (ntime, 2, 2) and list(sd["subgrid_id"]) == [1, 2]
        assert sd["time"][0] == 0.0 and sd["time"][-1] == model.t_end
        assert np.all(sd["subgrid_level"][0] == 0.01)
        i_change = 31  # 2020-02-01
        assert sd["subgrid_level"][i_change, 1, 0] - sd["subgrid_level"][i_change - 1, 1, 0] == \
            pytest.approx(1.0, rel=3.4e-4)  # `≈ 1.0f0`: the Basin level itself moves a little in a day
        basin_level = model.level()
        basin_level[1] += 1
        assert np.allclose(basin_level, sd["subgrid_level"][-1], rtol=1e-14)
        assert np.allclose(basin_level, model.subgrid_level(), rtol=1e-14)
        assert (tmp_path / "subgrid_level.nc").exists()
    finally:
        model.finalize()
    # BMI: update_subgrid_level fills basin.subgrid_level (core/test/bmi_test.jl:52-90, build/libribasim.jl:98-101)
    m2 = MODELS["two_basin"]()
    m2.solver.update(algorithm="ImplicitEuler", dt=3600.0)
    toml = m2.write(tmp_path / "model")
    b = RibasimBmi()
    b.initialize(toml)
    try:
        assert b.get_value_ptr("basin.subgrid_level").shape == (2,)
        b.update_until(86400.0)
        b.update_subgrid_level()
        assert np.allclose(b.get_value_ptr("basin.subgrid_level"), b.get_value_ptr("basin.level"), rtol=1e-14)
    finally:
        b.finalize()


def test_outlet_constraints():
    """core/test/run_models_test.jl:356-386: no Outlet flow while the upstream level is below the
    minimum upstream level; afterwards the Basin converges to the LevelBoundary level."""
    model = _model("outlet", n_members=1)
    try:
        res = model.run(saveat=86400.0)
        fd, bd = res.flow_data(), res.basin_data()
        level = model.p.nodes["level_boundary"]["level"][0]
        min_up = 2.0
        t_min = level.t[1] * (min_up - level.u[0]) / (level.u[1] - level.u[0])
        q = fd["flow_rate"][:, (fd["from_node_id"] == 2) & (fd["to_node_id"] == 3), 0].ravel()
        t_end_of_period = np.array(res.times[1:])
        assert np.all(q[t_end_of_period <= t_min] == 0.0)
        assert q.max() == pytest.approx(1e-3, rel=1e-6)
        assert model.level()[0, 0] == pytest.approx(level.u[2], abs=5e-2)
        assert model.status_counts["newton_failed"] == 0
    finally:
        model.finalize()


def test_flow_boundary_block_interpolation():
    """core/test/run_models_test.jl:729-761: with block interpolation the daily mean flows out of
    the (cyclic) FlowBoundary are the table's values, period after period."""
    model = _model("flow_boundary_interpolation", n_members=1)
    try:
        res = model.run(saveat=86400.0)
        fd = res.flow_data()
        q = fd["flow_rate"][:, fd["from_node_id"] == 1, 0].ravel()
        table = [1e-3, 2.5e-3, 0.0, 1e-3]
        assert np.allclose(q[0:4], table, rtol=1e-12, atol=1e-18)
        assert np.allclose(q[3:7], table, rtol=1e-12, atol=1e-18)
    finally:
        model.finalize()


def test_outlet_continuous_control_flow_split():
    """core/test/control_test.jl:263-288: two continuously controlled Outlets split the inflow
    through the LinearResistance 60 / 40 (mean flows per link from flow_data)."""
    model = _model("outlet_continuous_control", n_members=1)
    try:
        fd = model.run(saveat=86400.0).flow_data()

        def link_flow(a, b):
            return fd["flow_rate"][:, (fd["from_node_id"] == a) & (fd["to_node_id"] == b), 0].ravel()

        inflow = link_flow(2, 3)
        assert inflow.size == 366 and np.any(inflow > 0) and np.any(inflow < 0)
        scale = np.abs(inflow).max()
        for (a, b), frac in (((3, 4), 0.6), ((4, 6), 0.6), ((3, 5), 0.4), ((5, 7), 0.4)):
            # `≈` with rtol on the whole vector: norm-wise
            diff = link_flow(a, b) - np.maximum(frac * inflow, 0.0)
            assert np.linalg.norm(diff) <= 1e-3 * np.linalg.norm(np.maximum(frac * inflow, 0.0)) + 1e-12 * scale
    finally:
        model.finalize()


def test_time_dependent_flow_boundary():
    """core/test/time_test.jl:1-23: the daily mean flows out of the time-varying FlowBoundary
    follow 1 + sin^2 between March and October (rtol 0.005 on the vector)."""
    from datetime import timedelta

    model = _model("flow_boundary_time", dt=24 * 24 * 60.0, n_members=1)
    try:
        res = model.run(saveat=86400.0)
        fd = res.flow_data()
        start = model.tables.starttime
        summer = np.array([3 <= (start + timedelta(seconds=float(t))).month < 10 for t in fd["time"]])
        q = fd["flow_rate"][:, (fd["from_node_id"] == 1) & (fd["to_node_id"] == 2), 0].ravel()[summer]
        t = fd["time"][summer]
        expected = 1 + np.sin(0.5 * np.pi * (t - t[0]) / (t[-1] - t[0])) ** 2
        assert np.linalg.norm(q - expected) <= 0.005 * np.linalg.norm(expected)
    finally:
        model.finalize()


def test_vertical_flux_means():
    """core/test/time_test.jl:25-48 (basic_transient): mean evaporation ~ area(level) x potential
    evaporation (atol 1e-5: the area is that of the period's start), mean precipitation = fixed
    area x precipitation."""
    model = _model("basic_transient", n_members=1)
    try:
        res = model.run(saveat=86400.0)
        bd = res.basin_data()
        basin = model.p.basin
        om_area = None
        for i in range(len(basin.node_id)):
            level = bd["level"][:, i, 0]
            area = np.interp(level, basin.levels[i], basin.areas[i])
            pot_evap = np.array([basin.forcing["potential_evaporation"][i](float(t)) for t in bd["time"]])
            assert np.allclose(bd["evaporation"][:, i, 0], area * pot_evap, rtol=0, atol=1e-5)
            prec = np.array([basin.forcing["precipitation"][i](float(t)) for t in bd["time"]])
            assert np.allclose(bd["precipitation"][:, i, 0], basin.areas[i][-1] * prec, rtol=1e-8, atol=1e-18)
    finally:
        model.finalize()


def test_transient_pump_outlet():
    """core/test/time_test.jl:125-143: the Outlet fills the Basin up to the boundary level; the
    Pump's time-dependent flow rate is what flows over its inflow link in the second half year."""
    model = _model("transient_pump_outlet", n_members=1)
    try:
        S0 = model.storage()[0, 0]
        res = model.run(saveat=86400.0)
        assert S0 == pytest.approx(100.0, rel=1e-6)
        assert model.storage()[0, 0] == pytest.approx(110.0, rel=3.4e-4)  # `≈ 110.0f0`
        fd = res.flow_data()
        flow_4 = fd["flow_rate"][:, fd["link_id"] == 4, 0].ravel()
        assert np.allclose(flow_4[229:], 1e-5, rtol=1e-6, atol=0)
    finally:
        model.finalize()


def test_drought():
    """core/test/run_models_test.jl:706-727: Basin #2189 runs dry; after 100 days its storage,
    storage rate and infiltration are close to zero (the low-storage factor closes the sink)."""
    model = _model("drought", n_members=2)
    try:
        res = model.run(saveat=86400.0)
        bd = res.basin_data()
        b = list(bd["node_id"]).index(2189)
        late = bd["time"] > 100 * 86400.0
        assert np.all(np.abs(bd["storage"][late, b]) < 0.03)
        assert np.all(np.abs(bd["storage_rate"][late, b]) < 1e-8)
        assert np.all(np.abs(bd["infiltration"][late, b]) < 1e-8)
        assert model.status_counts["newton_failed"] == 0 and model.status_counts["negative_storage"] == 0
    finally:
        model.finalize()


def test_circular_flow_with_hysteresis_control(tmp_path):
    """core/test/control_test.jl:334-372, from the written control.nc and basin.nc (hourly saveat):
    the Pump is off initially (level <= 0.9), switches on only above 0.95 and off again only
    below 0.9.  The first days of the run hold the first three control records."""
    from scipy.io import netcdf_file

    model = _model("circular_flow", n_members=1)
    try:
        from ribasim_b200.results import Results
        res = Results(model)
        for k in range(1, 3 * 24 + 1):
            model.update_until(k * 3600.0)
            res.save()
        res.write(tmp_path)
    finally:
        model.finalize()
    with netcdf_file(str(tmp_path / "control.nc"), "r", mmap=False) as ds:
        control_times = ds.variables["time"][:].copy()
        truth = [b"".join(row).rstrip(b"\0").decode() for row in ds.variables["truth_state"][:]]
    with netcdf_file(str(tmp_path / "basin.nc"), "r", mmap=False) as ds:
        basin_times = ds.variables["time"][:].copy()
        node_ids = list(ds.variables["node_id"][:])
        level6 = ds.variables["level"][:][:, node_ids.index(6)].copy()
    assert len(truth) >= 3 and truth[:3] == ["F", "T", "F"]

    def level_at(t):
        return level6[np.argmax(basin_times >= t)]

    assert level_at(control_times[0]) <= 0.9 + 1e-10
    assert level_at(control_times[1]) > 0.95
    assert level_at(control_times[2]) <= 0.9 + 1e-2
