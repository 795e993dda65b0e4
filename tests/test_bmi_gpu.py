"""BMI surface on the GPU path, after core/test/bmi_test.jl and
python/ribasim_api/tests/test_bmi.py (fixed time stepping: the adaptive QNDF is out of scope)."""

from __future__ import annotations

import numpy as np
import pytest

from ribasim_b200.testmodels import MODELS

pytestmark = pytest.mark.gpu


@pytest.fixture
def basic_toml(tmp_path):
    m = MODELS["basic"]()
    m.solver.update(algorithm="ImplicitEuler", dt=3600.0)
    return m.write(tmp_path / "basic")


def _bmi(toml):
    from ribasim_b200.bmi import RibasimBmi
    b = RibasimBmi()
    b.initialize(toml)
    return b


def test_fixed_timestepping(basic_toml):
    # core/test/bmi_test.jl:30-50
    b = _bmi(basic_toml)
    try:
        dt = 3600.0
        assert b.get_time_units() == "s"
        assert b.get_time_step() == dt
        assert b.get_start_time() == 0.0 and b.get_current_time() == 0.0
        assert b.get_end_time() == pytest.approx(3.16224e7)
        b.update()
        assert b.get_current_time() == dt
        with pytest.raises(RuntimeError, match="already passed"):
            b.update_until(dt - 60)
        b.update_until(dt + 60)
        assert b.get_current_time() == dt + 60
        b.update()
        assert b.get_current_time() == 2 * dt + 60
        b.update_until(86400.0)
        assert b.get_current_time() == 86400.0
    finally:
        b.finalize()
    with pytest.raises(RuntimeError, match="Model not initialized"):
        b.update()  # core/test/libribasim_test.jl: error path


def test_get_value_ptr(basic_toml):
    # core/test/bmi_test.jl:52-90, python/ribasim_api/tests/test_bmi.py:95-110
    b = _bmi(basic_toml)
    try:
        storage0 = b.get_value_ptr("basin.storage")
        assert np.allclose(storage0, np.ones(4))
        assert np.allclose(b.get_value_ptr("basin.level"), 4 * [0.04471158417652035], atol=1e-6)
        with pytest.raises(KeyError, match="Unknown variable foo"):
            b.get_value_ptr("foo")
        names = ["basin.storage", "basin.level", "basin.infiltration", "basin.surface_runoff",
                 "basin.drainage", "basin.cumulative_infiltration",
                 "basin.cumulative_surface_runoff", "basin.cumulative_drainage",
                 "basin.subgrid_level", "user_demand.demand", "user_demand.cumulative_inflow"]
        first = {n: b.get_value_ptr(n) for n in names}
        b.update_until(86400.0)
        for n in names:  # get_value_ptr does not copy
            assert b.get_value_ptr(n) is first[n]
            assert b.get_var_type(n) == "double"
        assert not np.allclose(storage0, np.ones(4))
        assert b.get_component_name() == "Ribasim"
    finally:
        b.finalize()


def test_vertical_basin_flux(basic_toml):
    # core/test/bmi_test.jl:112-127: forcing written through the pointer is used by the solver
    b = _bmi(basic_toml)
    try:
        drainage = b.get_value_ptr("basin.drainage")
        flux = np.array([1.0, 2.0, 3.0, 4.0])
        drainage[:] = flux
        dt = 5 * 86400.0
        b.update_until(dt)
        assert np.allclose(b.get_value_ptr("basin.cumulative_drainage"), dt * flux, rtol=1e-12)
    finally:
        b.finalize()


def test_windowed_update_matches_stepwise(basic_toml):
    """update_until in one call (device windows) == hour by hour."""
    a, b = _bmi(basic_toml), _bmi(basic_toml)
    try:
        a.update_until(3 * 86400.0)
        for k in range(1, 73):
            b.update_until(k * 3600.0)
        assert np.array_equal(a.get_value_ptr("basin.storage"), b.get_value_ptr("basin.storage"))
    finally:
        a.finalize()
        b.finalize()


def test_stroboscopic_forcing(tmp_path):
    """core/test/run_models_test.jl:593-636: drainage and infiltration written through the BMI
    pointers on alternating days are integrated exactly into the cumulative arrays."""
    m = MODELS["bucket"]()
    m.solver.update(algorithm="ImplicitEuler", dt=3600.0)
    b = _bmi(m.write(tmp_path))
    try:
        drn, drn_sum = b.get_value_ptr("basin.drainage"), b.get_value_ptr("basin.cumulative_drainage")
        inf, inf_sum = b.get_value_ptr("basin.infiltration"), b.get_value_ptr("basin.cumulative_infiltration")
        nday, dt = 300, 86400.0
        inf_out, drn_out = np.full(nday, np.nan), np.full(nday, np.nan)
        for day in range(nday):
            if day % 2 == 0:
                drn[:] = 25.0 / dt
                inf[:] = 0.0
            else:
                drn[:] = 0.0
                inf[:] = 25.0 / dt
            b.update_until(day * dt)
            inf_out[day], drn_out[day] = inf_sum[0], drn_sum[0]
        d_drn, d_inf = np.diff(drn_out), np.diff(inf_out)
        # the forcing set before `update_until(day)` acts on the day that ends at `day`
        assert np.all(d_drn[0::2] == 0.0)
        assert np.allclose(d_drn[1::2], 25.0, rtol=0, atol=1e-10)
        assert np.allclose(d_inf[0::2], 25.0, rtol=0, atol=1e-10)
        assert np.all(d_inf[1::2] == 0.0)
    finally:
        b.finalize()



def test_user_demand_inflow(tmp_path):
    """core/test/bmi_test.jl:92-108 (minimal_subnetwork without allocation): the cumulative
    UserDemand inflows follow the demands, also after a demand is changed through the pointer."""
    m = MODELS["minimal_subnetwork"]()
    m.solver.update(algorithm="ImplicitEuler", dt=3600.0)
    b = _bmi(m.write(tmp_path))
    try:
        demand = b.get_value_ptr("user_demand.demand")
        inflow = b.get_value_ptr("user_demand.cumulative_inflow")
        day = 86400.0
        b.update_until(2 * day)
        assert inflow[0] == pytest.approx(2.0e-3 * day, abs=1e-2)
        # the block-interpolated demand of #6 doubles at day 1; implicit Euler evaluates the RHS at
        # the end of a step, so the fixed 1 h step that ends on the knot already sees the new
        # value (as the reference's ImplicitEuler would): one step of the demand difference on top
        # of the reference's adaptive-step tolerance
        assert inflow[1] == pytest.approx(3.0e-3 * day, abs=1e-3 * 3600.0 + 1e-2)
        assert inflow[1] >= 3.0e-3 * day - 1e-2
        demand[0] = 3.0e-3
        b.update_until(3 * day)
        assert inflow[0] == pytest.approx(5.0e-3 * day, abs=2e-3)
    finally:
        b.finalize()


def test_user_demand_demand_layout_is_node_fastest(tmp_path):
    """`user_demand.demand` is `vec()` of the reference's column-major [node, priority] matrix
    (core/src/bmi.jl:82-83; docs/dev/bmi.qmd): flat index = node + n_nodes * priority, and the array
    handed out aliases the model's demand matrix (a coupler writes through it)."""
    m = MODELS["minimal_subnetwork"]()
    m.solver.update(algorithm="ImplicitEuler", dt=3600.0)
    # second demand priority on node 5 only: the matrix becomes 2 nodes x 2 priorities
    m.tables["UserDemand / static"].append(
        dict(node_id=5, demand=4e-3, return_factor=0.9, min_level=0.9, demand_priority=2))
    b = _bmi(m.write(tmp_path))
    try:
        flat = b.get_value_ptr("user_demand.demand")
        mat = b.model.p.nodes["user_demand"]["demand"]
        assert mat.shape == (2, 2) and flat.shape == (4,)
        assert np.shares_memory(flat, mat)
        n_nodes = 2
        assert flat[0 + n_nodes * 0] == 1e-3 and flat[0 + n_nodes * 1] == 4e-3   # node 5
        assert flat[1 + n_nodes * 1] == 0.0                                      # node 6 has no priority 2
        flat[0 + n_nodes * 1] = 7e-3
        assert mat[0, 1] == 7e-3
        inflow = b.get_value_ptr("user_demand.cumulative_inflow")
        b.update_until(86400.0)
        # node 5 now asks for 1e-3 + 7e-3 (the sum over its priorities, solve.jl:348-404), limited by
        # what the pump delivers into its basin (4e-3) minus what node 6 takes
        assert inflow[0] > 1.5e-3 * 86400.0
    finally:
        b.finalize()
