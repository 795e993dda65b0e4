"""Test models and synthetic network generators, as plain tables.

The named models re-express the reference's `ribasim_testmodels` definitions (only the tables
the core reads; geometry, subgrid and concentration tables are not needed on the hot path):

    basic                      python/ribasim_testmodels/ribasim_testmodels/basic.py:26-222
    basic_transient            basic.py:225-283
    backwater                  backwater.py:16-84
    trivial                    trivial.py:10-62
    linear_resistance ... pid_control_equation   equations.py
    pid_control                pid_control.py:19-63
    outlet_continuous_control  continuous_control.py:18-86
    user_demand                allocation.py:30-95
    pump_discrete_control      discrete_control.py:23-126
    flow_condition, level_boundary_condition, tabulated_rating_curve_control,
    compound_variable_condition, storage_condition, connector_node_flow_condition,
    transient_condition, circular_flow  discrete_control.py:128-938
    discrete_control_of_pid_control  pid_control.py:66-152
    outlet, cyclic_time, drought, flow_boundary_interpolation  basic.py:364-627
    flow_boundary_time, transient_pump_outlet  time.py:12-96
    two_basin                  two_basin.py:10-79

The synthetic generators implement the configurations fixed in SURVEY.md §8(d).
"""

from __future__ import annotations

from datetime import datetime, timedelta

import numpy as np

from .tables import TIME_FMT, ModelTables


def _model(end=datetime(2021, 1, 1), **kw) -> ModelTables:
    return ModelTables(starttime=datetime(2020, 1, 1), endtime=end, **kw)


# ----------------------------------------------------------------------------- reference models
def basic_model() -> ModelTables:
    m = _model()
    for node_id in (1, 3, 6, 9):
        m.add_node(
            "Basin", node_id,
            profile=dict(area=[0.01, 1000.0], level=[0.0, 1.0]),
            static=dict(potential_evaporation=0.001 / 86400, precipitation=0.001 / 86400,
                        surface_runoff=1 / 86400),
            state=dict(level=0.04471158417652035),
        )
    m.add_node("LinearResistance", 12, static=dict(resistance=5e3))
    m.add_node("LinearResistance", 10, static=dict(resistance=(3600.0 * 24) / 100.0))
    m.add_node("ManningResistance", 2, static=dict(length=900.0, manning_n=0.04,
                                                   profile_width=1.0, profile_slope=3.0))
    q = 10 / 86400
    m.add_node("TabulatedRatingCurve", 8, static=dict(level=[0.0, 1.0], flow_rate=[0.0, 0.6 * q]))
    m.add_node("TabulatedRatingCurve", 5, static=dict(level=[0.0, 1.0], flow_rate=[0.0, 0.3 * q]))
    m.add_node("TabulatedRatingCurve", 4, static=dict(level=[0.0, 1.0], flow_rate=[0.0, 0.1 * q]))
    m.add_node("Pump", 7, static=dict(flow_rate=0.5 / 3600))
    m.add_node("FlowBoundary", 15, static=dict(flow_rate=1e-4))
    m.add_node("FlowBoundary", 16, static=dict(flow_rate=1e-4))
    m.add_node("LevelBoundary", 11, static=dict(level=1.0))
    m.add_node("LevelBoundary", 17, static=dict(level=1.5))
    m.add_node("Junction", 13)
    m.add_node("Terminal", 14)
    for a, b in [(1, 2), (2, 3), (3, 8), (3, 5), (3, 4), (5, 6), (6, 7), (8, 13), (7, 13),
                 (13, 9), (9, 10), (11, 12), (12, 3), (4, 14), (15, 6), (16, 1), (10, 17)]:
        m.add_link(a, b)
    return m


def basic_transient_model() -> ModelTables:
    m = basic_model()
    days = [m.starttime + timedelta(days=k) for k in range(367)]
    doy = np.array([d.timetuple().tm_yday for d in days], dtype=float)
    evaporation = (-1.0 * np.cos(doy / 365.0 * 2 * np.pi) + 1.0) * 0.0025 / 86400
    rng = np.random.default_rng(seed=0)
    precipitation = rng.lognormal(mean=-1.0, sigma=1.7, size=len(days)) * 0.001 / 86400
    del m.tables["Basin / static"]
    rows = []
    for node_id in (1, 3, 6, 9):
        for d, e, pr in zip(days, evaporation, precipitation):
            rows.append(dict(node_id=node_id, time=d.strftime(TIME_FMT), drainage=0.0,
                             potential_evaporation=float(e), infiltration=0.0,
                             precipitation=float(pr)))
    m.tables["Basin / time"] = rows
    m.tables["Basin / state"] = [dict(node_id=i, level=1.4) for i in (1, 3, 6, 9)]
    return m


def flow_boundary_time_model() -> ModelTables:
    """One Basin fed by a constant and a time-varying (linearly interpolated) FlowBoundary
    (python/ribasim_testmodels/ribasim_testmodels/time.py:12-56)."""
    m = _model(interpolation=dict(flow_boundary="linear"))
    m.add_node("FlowBoundary", 3, static=dict(flow_rate=1.0))
    n_times = 100
    t0, t1 = datetime(2020, 3, 1), datetime(2020, 10, 1)
    span = (t1 - t0).total_seconds()
    times = [t0 + timedelta(seconds=round(span * k / (n_times - 1))) for k in range(n_times)]
    flow_rate = 1 + np.sin(np.pi * np.linspace(0, 0.5, n_times)) ** 2
    m.add_node("FlowBoundary", 1, time=dict(time=[t.strftime(TIME_FMT) for t in times],
                                            flow_rate=[float(q) for q in flow_rate]))
    m.add_node("Basin", 2, profile=dict(area=[0.01, 1000.0], level=[0.0, 1.0]),
               state=dict(level=0.04471158417652035))
    m.add_link(1, 2)
    m.add_link(3, 2)
    return m


def two_basin_model() -> ModelTables:
    """Two unconnected Basins with subgrid h(h) relations, one of them time dependent
    (python/ribasim_testmodels/ribasim_testmodels/two_basin.py:10-79)."""
    m = _model()
    m.add_node("FlowBoundary", 1, static=dict(flow_rate=1e-2))
    shared = dict(profile=dict(area=[400.0, 400.0], level=[0.0, 1.0]), state=dict(level=0.01))
    m.add_node("Basin", 2, **shared,
               subgrid=dict(subgrid_id=1, basin_level=[0.0, 1.0], subgrid_level=[0.0, 1.0]))
    m.add_node("Basin", 3, **shared,
               subgrid_time=dict(subgrid_id=2,
                                 time=["2020-01-01 00:00:00", "2020-01-01 00:00:00",
                                       "2020-02-01 00:00:00", "2020-02-01 00:00:00"],
                                 basin_level=[0.0, 1.0, 0.0, 1.0], subgrid_level=[0.0, 1.0, 1.0, 2.0]))
    m.add_node("TabulatedRatingCurve", 4, static=dict(level=[0.0, 1.0], flow_rate=[0.0, 0.01]))
    m.add_node("Terminal", 5)
    for a, b in [(1, 2), (3, 4), (4, 5)]:
        m.add_link(a, b)
    return m


def backwater_model(n_basins: int = 50) -> ModelTables:
    """FlowBoundary -> n x (Basin, ManningResistance) -> huge Basin (backwater.py:16-84).

    The shipped model has n_basins = 50 (+ the downstream 1e10 m2 Basin)."""
    m = _model()
    m.add_node("FlowBoundary", 1, static=dict(flow_rate=5.0))
    for k in range(n_basins):
        bid = 2 + 2 * k
        m.add_node("Basin", bid, profile=dict(area=[20.0, 20.0], level=[0.0, 1.0]),
                   state=dict(level=0.05))
        m.add_node("ManningResistance", bid + 1,
                   static=dict(length=20.0, manning_n=0.04, profile_width=1.0, profile_slope=0.0))
        m.add_link(1 if k == 0 else bid - 1, bid)
        m.add_link(bid, bid + 1)
    last = 2 + 2 * n_basins
    m.add_node("Basin", last, state=dict(level=2.0),
               profile=dict(level=[0.0, 1.0], area=[1e10, 1e10]))
    m.add_link(last - 1, last)
    return m


def trivial_model() -> ModelTables:
    m = _model()
    m.add_node("Basin", 6,
               static=dict(precipitation=0.002 / 86400, potential_evaporation=0.001 / 86400),
               profile=dict(area=[0.01, 1000.0], level=[0.0, 1.0]),
               state=dict(level=0.04471158417652035))
    m.add_node("Terminal", 2147483647)
    m.add_node("TabulatedRatingCurve", 0,
               static=dict(level=[0.0, 1.0], flow_rate=[0.0, 10 / 86400]))
    m.add_link(6, 0, link_id=100)
    m.add_link(0, 2147483647)
    return m


def bucket_model() -> ModelTables:
    m = _model()
    m.add_node("Basin", 1, profile=dict(area=[1000.0, 1000.0], level=[0.0, 1.0]),
               state=dict(level=1.0),
               static=dict(precipitation=0.0, potential_evaporation=0.0, drainage=0.0,
                           infiltration=0.0))
    return m


def leaky_bucket_model() -> ModelTables:
    """One Basin with daily forcing that has missing values: a missing value keeps the previous one
    (python/ribasim_testmodels/ribasim_testmodels/bucket.py:42-70, core/src/callback.jl:826-838)."""
    m = _model(end=datetime(2020, 1, 5))
    nan = float("nan")
    m.add_node("Basin", 1, profile=dict(area=[1000.0, 1000.0], level=[0.0, 1.0]), state=dict(level=1.0))
    drainage = [0.003, nan, 0.001, 0.002, 0.0]
    infiltration = [nan, 0.001, 0.002, 0.0, 0.0]
    m.tables["Basin / time"] = [
        dict(node_id=1, time=(m.starttime + timedelta(days=k)).strftime(TIME_FMT), drainage=drainage[k],
             potential_evaporation=nan, infiltration=infiltration[k], precipitation=nan)
        for k in range(5)]
    return m


def linear_resistance_model() -> ModelTables:
    m = _model()
    m.add_node("Basin", 1, profile=dict(area=[100.0, 100.0], level=[0.0, 10.0]),
               state=dict(level=10.0))
    m.add_node("LinearResistance", 2, static=dict(resistance=5e4, max_flow_rate=6e-5))
    m.add_node("LevelBoundary", 3, static=dict(level=5.0))
    m.add_link(1, 2)
    m.add_link(2, 3)
    return m


def rating_curve_model() -> ModelTables:
    m = _model()
    m.add_node("Basin", 1, profile=dict(area=[0.01, 100.0, 100.0], level=[0.0, 1.0, 12.0]),
               state=dict(level=10.5))
    level = np.linspace(1, 12, 100)
    flow_rate = np.square(level - 1.0) / 86400
    m.add_node("TabulatedRatingCurve", 2, static=dict(level=level, flow_rate=flow_rate))
    m.add_node("Terminal", 3)
    m.add_link(1, 2)
    m.add_link(2, 3)
    return m


def rating_curve_between_basins_model() -> ModelTables:
    m = _model()
    level = np.linspace(1, 12, 100)
    flow_rate = np.square(level - 1.0) / 86400
    m.add_node("Basin", 1, profile=dict(area=[0.01, 50.0, 100.0], level=[0.0, 1.0, 12.0]),
               state=dict(level=10.5))
    m.add_node("TabulatedRatingCurve", 2, static=dict(level=level, flow_rate=flow_rate))
    m.add_node("Basin", 3, profile=dict(area=[0.01, 50.0, 100.0], level=[0.0, 1.0, 12.0]),
               state=dict(level=5.5))
    m.add_node("TabulatedRatingCurve", 4, static=dict(level=level, flow_rate=flow_rate))
    m.add_node("Terminal", 5)
    for a, b in [(1, 2), (2, 3), (3, 4), (4, 5)]:
        m.add_link(a, b)
    return m


def manning_resistance_model() -> ModelTables:
    m = _model()
    prof = dict(area=[0.01, 100.0, 100.0], level=[0.0, 1.0, 10.0])
    m.add_node("Basin", 1, profile=prof, state=dict(level=9.5))
    m.add_node("ManningResistance", 2, static=dict(manning_n=1e7, profile_width=50.0,
                                                   profile_slope=0.0, length=2000.0))
    m.add_node("Basin", 3, profile=prof, state=dict(level=4.5))
    m.add_link(1, 2)
    m.add_link(2, 3)
    return m


def misc_nodes_model() -> ModelTables:
    m = _model(solver=dict(dt=86400.0, algorithm="Euler"))
    prof = dict(area=[0.01, 100.0, 100.0], level=[0.0, 1.0, 2.0])
    m.add_node("FlowBoundary", 1, static=dict(flow_rate=1.5e-4))
    m.add_node("Basin", 3, profile=prof, state=dict(level=10.5))
    m.add_node("Pump", 4, static=dict(flow_rate=1e-4))
    m.add_node("Basin", 5, profile=prof, state=dict(level=10.5))
    for a, b in [(1, 3), (3, 4), (4, 5)]:
        m.add_link(a, b)
    return m


def pid_control_equation_model() -> ModelTables:
    m = _model(end=datetime(2020, 1, 1, 0, 5, 0))
    m.add_node("Basin", 1, profile=dict(area=[0.01, 100.0, 100.0], level=[0.0, 1.0, 2.0]),
               state=dict(level=10.5))
    m.add_node("Pump", 2, static=dict(flow_rate=0.0))
    m.add_node("Terminal", 3)
    m.add_node("PidControl", 4, static=dict(listen_node_id=1, target=10.0, proportional=-2.5,
                                            integral=-0.001, derivative=10.0))
    m.add_link(1, 2)
    m.add_link(2, 3)
    m.add_link(4, 2)
    return m


def pid_control_model() -> ModelTables:
    m = _model(end=datetime(2021, 12, 1))
    m.add_node("FlowBoundary", 1, static=dict(flow_rate=1e-3))
    m.add_node("Basin", 2, profile=dict(area=[1000.0, 1000.0], level=[0.0, 1.0]),
               state=dict(level=6.0))
    m.add_node("Pump", 3, static=dict(flow_rate=0.0))
    m.add_node("LevelBoundary", 4, static=dict(level=5.5))
    m.add_node("PidControl", 5, time=dict(
        time=["2020-01-01 00:00:00", "2020-12-01 00:00:00", "2021-12-01 00:00:00"],
        listen_node_id=2, target=[5.5, 6.5, 6.5], proportional=-5e-3, integral=-7e-10,
        derivative=-8e3))
    for a, b in [(1, 2), (2, 3), (3, 4), (5, 3)]:
        m.add_link(a, b)
    return m


def outlet_continuous_control_model() -> ModelTables:
    m = _model()
    t0, t1 = datetime(2020, 1, 1), datetime(2021, 1, 1)
    total_ms = int((t1 - t0).total_seconds() * 1000)
    # pd.date_range(periods=100, unit="ms") truncates to whole milliseconds; the core then
    # reads whole seconds (time strings are written with second resolution)
    times = [t0 + timedelta(milliseconds=int(total_ms * k / 99)) for k in range(100)]
    level = 6.0 + np.sin(np.linspace(0, 6 * np.pi, 100))
    m.add_node("LevelBoundary", 1,
               time=dict(time=[t.strftime(TIME_FMT) for t in times], level=level))
    m.add_node("LinearResistance", 2, static=dict(resistance=10.0))
    m.add_node("Basin", 3, profile=dict(area=[10000.0, 10000.0], level=[0.0, 1.0]),
               state=dict(level=10.0))
    m.add_node("Outlet", 4, static=dict(flow_rate=1.0))
    m.add_node("Outlet", 5, static=dict(flow_rate=1.0))
    m.add_node("Terminal", 6)
    m.add_node("Terminal", 7)
    m.add_node("ContinuousControl", 8,
               variable=dict(listen_node_id=2, variable="flow_rate"),
               function=dict(input=[0.0, 1.0], output=[0.0, 0.6],
                             controlled_variable="flow_rate"))
    m.add_node("ContinuousControl", 9,
               variable=dict(listen_node_id=2, variable="flow_rate"),
               function=dict(input=[0.0, 1.0], output=[0.0, 0.4],
                             controlled_variable="flow_rate"))
    for a, b in [(1, 2), (2, 3), (3, 4), (3, 5), (4, 6), (5, 7), (8, 4), (9, 5)]:
        m.add_link(a, b)
    return m


def user_demand_model() -> ModelTables:
    m = _model(solver=dict(algorithm="Tsit5"))
    m.add_node("Basin", 1, profile=dict(area=[1000.0, 1000.0], level=[0.0, 2.0]),
               state=dict(level=1.0))
    m.add_node("UserDemand", 2, static=dict(demand=1e-4, return_factor=0.9, min_level=0.9,
                                            demand_priority=1))
    m.add_node("UserDemand", 3, time=dict(
        time=["2020-06-01 00:00:00", "2020-06-01 01:00:00", "2020-07-01 00:00:00",
              "2020-07-01 01:00:00"],
        demand=[0.0, 3e-4, 3e-4, 0.0], return_factor=0.4, min_level=0.5, demand_priority=1))
    days = [datetime(2020, 8, 1) + timedelta(days=k) for k in range(32)] + \
           [datetime(2020, 10, 1) + timedelta(days=k) for k in range(32)]
    m.add_node("UserDemand", 4, time=dict(
        time=[d.strftime(TIME_FMT) for d in days],
        demand=np.concatenate([np.linspace(0.0, 1e-4, num=32), np.linspace(2e-4, 0.0, num=32)]),
        min_level=0.0,
        return_factor=np.concatenate([np.linspace(0.0, 0.1, num=32),
                                      np.linspace(0.2, 0.3, num=32)]),
        demand_priority=1))
    m.add_node("Terminal", 5)
    for a, b in [(1, 2), (1, 3), (1, 4), (2, 5), (3, 5), (4, 5)]:
        m.add_link(a, b)
    return m


def minimal_subnetwork_model() -> ModelTables:
    """FlowBoundary -> Basin -> Pump -> Basin with two UserDemand nodes, run without allocation
    (python/ribasim_testmodels/ribasim_testmodels/allocation.py:187-241)."""
    m = _model()
    shared = dict(profile=dict(area=[1000.0, 1000.0], level=[0.0, 1000.0]), state=dict(level=1.0))
    m.add_node("FlowBoundary", 1, static=dict(flow_rate=2.0e-3))
    m.add_node("Basin", 2, **shared)
    m.add_node("Pump", 3, static=dict(flow_rate=4e-3, max_flow_rate=4e-3))
    m.add_node("Basin", 4, **shared)
    m.add_node("UserDemand", 5, static=dict(demand=1e-3, return_factor=0.9, min_level=0.9, demand_priority=1))
    m.add_node("UserDemand", 6, time=dict(time=["2020-01-01 00:00:00", "2020-01-02 00:00:00"],
                                          demand=[1e-3, 2e-3], return_factor=0.9, min_level=0.9,
                                          demand_priority=1))
    for a, b in [(1, 2), (2, 3), (3, 4), (4, 5), (4, 6), (5, 4), (6, 4)]:
        m.add_link(a, b)
    return m


def basic_basin_model(profile: str = "storage") -> ModelTables:
    """Two monthly FlowBoundary series into a small parabolic Basin that drains over a weir; the
    profile is given by area, by storage or by both (basic.py:630-760:
    basic_basin_only_area / basic_basin_only_storage / basic_basin_both_area_and_storage)."""
    m = ModelTables(starttime=datetime(2022, 1, 1), endtime=datetime(2023, 1, 1))
    months = [datetime(2022 + (k // 12), k % 12 + 1, 1).strftime(TIME_FMT) for k in range(13)]
    main = [74.7, 57.9, 63.2, 183.9, 91.8, 47.5, 32.6, 27.6, 26.5, 25.1, 39.3, 37.8, 57.9]
    minor = [16.3, 3.8, 3.0, 37.6, 18.2, 11.1, 12.9, 12.2, 11.2, 10.8, 15.1, 14.3, 11.8]
    m.add_node("FlowBoundary", 1, time=dict(time=months, flow_rate=main))
    m.add_node("FlowBoundary", 2, time=dict(time=months, flow_rate=minor))
    levels = [0.0, 1.0, 2.0, 3.0, 4.0, 5.0]
    cols: dict = dict(level=levels)
    if profile in ("area", "both"):
        cols["area"] = [(h + 1) * np.pi for h in levels]
    if profile in ("storage", "both"):
        cols["storage"] = [np.pi / 2 * ((h + 1) ** 2 - 1) for h in levels]
    m.add_node("Basin", 3, profile=cols, state=dict(level=4.0))
    m.add_node("TabulatedRatingCurve", 4, static=dict(level=[0.0, 2.0, 5.0], flow_rate=[0.0, 50.0, 200.0]))
    m.add_node("Terminal", 5)
    for a, b in [(1, 3), (2, 3), (3, 4), (4, 5)]:
        m.add_link(a, b)
    return m


def pump_discrete_control_model() -> ModelTables:
    m = _model()
    m.add_node("Basin", 1, state=dict(level=1.0),
               profile=dict(level=[0.0, 1.0], area=[100.0, 100.0]))
    m.add_node("LinearResistance", 2, static=dict(resistance=[1e5, float("inf")],
                                                  control_state=["active", "inactive"]))
    m.add_node("Basin", 3, state=dict(level=1e-5), static=dict(precipitation=1e-9),
               profile=dict(level=[0.0, 1.0], area=[100.0, 100.0]))
    m.add_node("Pump", 4, static=dict(flow_rate=[0.0, 1e-5], control_state=["off", "on"]))
    m.add_node("DiscreteControl", 5,
               variable=dict(listen_node_id=[1, 3], variable="level", compound_variable_id=[1, 2]),
               condition=dict(threshold_high=[0.8, 0.4], compound_variable_id=[1, 2],
                              condition_id=[1, 1]),
               logic=dict(truth_state=["FF", "TF", "FT", "TT"],
                          control_state=["on", "off", "off", "on"]))
    m.add_node("DiscreteControl", 6,
               variable=dict(listen_node_id=3, variable="level", compound_variable_id=1),
               condition=dict(threshold_high=0.45, compound_variable_id=1, condition_id=1),
               logic=dict(truth_state=["T", "F"], control_state=["inactive", "active"]))
    for a, b in [(1, 2), (2, 3), (1, 4), (4, 3), (5, 4), (6, 2)]:
        m.add_link(a, b)
    return m


def discrete_control_of_pid_control_model() -> ModelTables:
    """A DiscreteControl node listening to a (transient) LevelBoundary sets the target level of a
    PidControl node (python/ribasim_testmodels/ribasim_testmodels/pid_control.py:66-152)."""
    m = _model(end=datetime(2020, 12, 1))
    m.add_node("LevelBoundary", 1, time=dict(
        time=["2020-01-01 00:00:00", "2020-07-01 00:00:00", "2021-01-01 00:00:00"], level=[7.0, 3.0, 3.0]))
    m.add_node("Outlet", 2, static=dict(flow_rate=0.0))
    m.add_node("Basin", 3, state=dict(level=6.0), profile=dict(area=[1000.0, 1000.0], level=[0.0, 1.0]))
    m.add_node("TabulatedRatingCurve", 4, static=dict(level=[0.0, 1.0], flow_rate=[0.0, 10 / 86400]))
    m.add_node("Terminal", 5)
    m.add_node("PidControl", 6, static=dict(
        listen_node_id=3, control_state=["target_high", "target_low"], target=[5.0, 3.0],
        proportional=1e-2, integral=1e-8, derivative=-1e-1))
    m.add_node("DiscreteControl", 7,
               variable=dict(listen_node_id=1, variable="level", compound_variable_id=1),
               condition=dict(threshold_high=5.0, compound_variable_id=1, condition_id=1),
               logic=dict(truth_state=["T", "F"], control_state=["target_high", "target_low"]))
    for a, b in [(1, 2), (2, 3), (3, 4), (4, 5), (6, 2), (7, 6)]:
        m.add_link(a, b)
    return m


def manning_discrete_control_model() -> ModelTables:
    """Two Basins joined by a ManningResistance whose roughness and width a DiscreteControl node
    switches on the upstream level (ManningResistance control_state rows, read.jl:328-427)."""
    m = _model(end=datetime(2020, 3, 1))
    m.add_node("FlowBoundary", 1, static=dict(flow_rate=3e-3))
    m.add_node("Basin", 2, state=dict(level=0.3), profile=dict(area=[100.0, 100.0], level=[0.0, 1.0]))
    m.add_node("ManningResistance", 3, static=dict(
        length=[100.0, 80.0], manning_n=[0.08, 0.02], profile_width=[1.0, 2.0], profile_slope=[1.0, 0.5],
        control_state=["rough", "smooth"]))
    m.add_node("Basin", 4, state=dict(level=0.1), profile=dict(area=[100.0, 100.0], level=[0.0, 1.0]))
    m.add_node("TabulatedRatingCurve", 5, static=dict(level=[0.0, 1.0], flow_rate=[0.0, 5e-3]))
    m.add_node("Terminal", 6)
    m.add_node("DiscreteControl", 7,
               variable=dict(listen_node_id=2, variable="level", compound_variable_id=1),
               condition=dict(threshold_high=0.6, threshold_low=0.4, compound_variable_id=1, condition_id=1),
               logic=dict(truth_state=["T", "F"], control_state=["smooth", "rough"]))
    for a, b in [(1, 2), (2, 3), (3, 4), (4, 5), (5, 6), (7, 3)]:
        m.add_link(a, b)
    return m


def flow_condition_model() -> ModelTables:
    """DiscreteControl on the flow rate of a FlowBoundary, 60 days ahead
    (python/ribasim_testmodels/ribasim_testmodels/discrete_control.py:128-191)."""
    m = _model(interpolation=dict(flow_boundary="linear"))
    m.add_node("FlowBoundary", 1, time=dict(time=["2020-01-01 00:00:00", "2022-01-01 00:00:00"],
                                            flow_rate=[0.0, 40 / 86400]))
    m.add_node("Basin", 2, profile=dict(level=[0.0, 1.0], area=[100.0, 100.0]), state=dict(level=2.5))
    m.add_node("Pump", 3, static=dict(flow_rate=[0.0, 1e-3], control_state=["off", "on"]))
    m.add_node("Terminal", 4)
    m.add_node("DiscreteControl", 5,
               variable=dict(listen_node_id=1, variable="flow_rate", look_ahead=60 * 86400.0,
                             compound_variable_id=1),
               condition=dict(threshold_high=20 / 86400, compound_variable_id=1, condition_id=1),
               logic=dict(truth_state=["T", "F"], control_state=["off", "on"]))
    for a, b in [(1, 2), (2, 3), (3, 4), (5, 3)]:
        m.add_link(a, b)
    return m


def level_boundary_condition_model() -> ModelTables:
    """DiscreteControl on the level of a transient LevelBoundary, 60 days ahead
    (python/ribasim_testmodels/ribasim_testmodels/discrete_control.py:194-264)."""
    m = _model()
    m.add_node("LevelBoundary", 1, time=dict(
        time=["2020-01-01 00:00:00", "2020-06-01 00:00:00", "2022-01-01 00:00:00"], level=[5.0, 7.0, 10.0]))
    m.add_node("LinearResistance", 2, static=dict(resistance=5e3))
    m.add_node("Basin", 3, profile=dict(level=[0.0, 1.0], area=[100.0, 100.0]), state=dict(level=2.5))
    m.add_node("Outlet", 4, static=dict(flow_rate=[0.5 / 3600, 0.0], control_state=["on", "off"]))
    m.add_node("Terminal", 5)
    m.add_node("DiscreteControl", 6,
               variable=dict(listen_node_id=1, variable="level", look_ahead=60 * 86400.0,
                             compound_variable_id=1),
               condition=dict(threshold_high=6.0, compound_variable_id=1, condition_id=1),
               logic=dict(truth_state=["T", "F"], control_state=["on", "off"]))
    for a, b in [(1, 2), (2, 3), (3, 4), (4, 5), (6, 4)]:
        m.add_link(a, b)
    return m


def tabulated_rating_curve_control_model() -> ModelTables:
    """DiscreteControl raises the crest of a rating curve above a threshold level
    (python/ribasim_testmodels/ribasim_testmodels/discrete_control.py:268-338)."""
    m = _model()
    m.add_node("Basin", 1, static=dict(precipitation=0.002 / 86400),
               state=dict(level=0.04471158417652035), profile=dict(area=[0.01, 1000.0], level=[0.0, 1.0]))
    m.add_node("TabulatedRatingCurve", 2, static=dict(
        level=[0.0, 1.2, 0.0, 1.0], flow_rate=[0.0, 1 / 86400, 0.0, 1 / 86400],
        control_state=["low", "low", "high", "high"]))
    m.add_node("Terminal", 3)
    m.add_node("DiscreteControl", 4,
               variable=dict(listen_node_id=1, variable="level", compound_variable_id=1),
               condition=dict(threshold_high=0.5, compound_variable_id=1, condition_id=1),
               logic=dict(truth_state=["T", "F"], control_state=["low", "high"]))
    for a, b in [(1, 2), (2, 3), (4, 2)]:
        m.add_link(a, b)
    return m


def compound_variable_condition_model() -> ModelTables:
    """A condition on the weighted sum of two FlowBoundary rates
    (python/ribasim_testmodels/ribasim_testmodels/discrete_control.py:341-397)."""
    m = _model(interpolation=dict(flow_boundary="linear"))
    m.add_node("Basin", 1, profile=dict(area=[1000.0, 1000.0], level=[0.0, 1.0]), state=dict(level=1.0))
    m.add_node("FlowBoundary", 2, static=dict(flow_rate=0.0))
    m.add_node("FlowBoundary", 3, time=dict(time=["2020-01-01 00:00:00", "2021-01-01 00:00:00"],
                                            flow_rate=[0.0, 2.0]))
    m.add_node("Pump", 4, static=dict(control_state=["Off", "On"], flow_rate=[0.0, 1.0]))
    m.add_node("Terminal", 5)
    m.add_node("DiscreteControl", 6,
               variable=dict(listen_node_id=[2, 3], variable="flow_rate", weight=0.5, compound_variable_id=1),
               condition=dict(threshold_high=0.5, compound_variable_id=1, condition_id=1),
               logic=dict(truth_state=["T", "F"], control_state=["On", "Off"]))
    for a, b in [(2, 1), (3, 1), (1, 4), (4, 5), (6, 4)]:
        m.add_link(a, b)
    return m


def connector_node_flow_condition_model() -> ModelTables:
    """A condition on the flow through a LinearResistance that the control node itself closes
    (python/ribasim_testmodels/ribasim_testmodels/discrete_control.py:540-590)."""
    m = _model()
    m.add_node("Basin", 1, profile=dict(area=[1000.0, 1000.0], level=[0.0, 1.0]), state=dict(level=20.0))
    m.add_node("LinearResistance", 2, static=dict(control_state=["On", "Off"], resistance=[1e4, float("inf")]))
    m.add_node("Basin", 3, profile=dict(area=[1000.0, 1000.0], level=[0.0, 1.0]), state=dict(level=10.0))
    m.add_node("DiscreteControl", 4,
               variable=dict(listen_node_id=2, variable="flow_rate", compound_variable_id=1),
               condition=dict(threshold_high=1e-4, compound_variable_id=1, condition_id=1),
               logic=dict(truth_state=["T", "F"], control_state=["On", "Off"]))
    for a, b in [(1, 2), (2, 3), (4, 2)]:
        m.add_link(a, b)
    return m


def storage_condition_model() -> ModelTables:
    """A Pump switched on the storage of a Basin
    (python/ribasim_testmodels/ribasim_testmodels/discrete_control.py:495-537)."""
    m = _model()
    m.add_node("FlowBoundary", 1, static=dict(flow_rate=1e-3))
    m.add_node("Basin", 2, profile=dict(area=[1000.0, 1000.0], level=[0.0, 10.0]), state=dict(level=5.0))
    m.add_node("Pump", 3, static=dict(control_state=["off", "on"], flow_rate=[0.0, 1e-3]))
    m.add_node("Terminal", 4)
    m.add_node("DiscreteControl", 5,
               variable=dict(compound_variable_id=1, listen_node_id=2, variable="storage"),
               condition=dict(compound_variable_id=1, condition_id=1, threshold_high=7500.0),
               logic=dict(truth_state=["F", "T"], control_state=["off", "on"]))
    for a, b in [(1, 2), (2, 3), (3, 4), (5, 3)]:
        m.add_link(a, b)
    return m


def transient_condition_model() -> ModelTables:
    """A threshold that changes in time (cyclic)
    (python/ribasim_testmodels/ribasim_testmodels/discrete_control.py:755-798)."""
    m = _model(end=datetime(2020, 3, 1))
    m.add_node("LevelBoundary", 1, static=dict(level=2.0))
    m.add_node("Pump", 2, static=dict(control_state=["A", "B"], flow_rate=[1.0, 2.0]))
    m.add_node("Basin", 3, state=dict(level=2.0), profile=dict(level=[0.0, 1.0], area=[100.0, 100.0]))
    m.add_node("DiscreteControl", 4, cyclic_time=True,
               variable=dict(listen_node_id=1, variable="level", compound_variable_id=1),
               condition=dict(compound_variable_id=1, condition_id=1, threshold_high=[1.0, 3.0, 1.0],
                              time=["2020-01-01 00:00:00", "2020-02-01 00:00:00", "2020-03-01 00:00:00"]),
               logic=dict(truth_state=["F", "T"], control_state=["A", "B"]))
    for a, b in [(1, 2), (2, 3), (4, 2)]:
        m.add_link(a, b)
    return m


def outlet_model() -> ModelTables:
    """An Outlet between a transient LevelBoundary and a Basin that meets its constraints
    (python/ribasim_testmodels/ribasim_testmodels/basic.py:364-407)."""
    m = _model()
    m.add_node("Basin", 3, profile=dict(area=[1000.0, 1000.0], level=[0.0, 10.0]), state=dict(level=0.0))
    m.add_node("LevelBoundary", 1, time=dict(
        time=["2020-01-01 00:00:00", "2020-06-01 00:00:00", "2021-01-01 00:00:00"], level=[1.0, 3.0, 3.0]))
    m.add_node("Outlet", 2, static=dict(flow_rate=1e-3, min_upstream_level=2.0))
    m.add_link(1, 2)
    m.add_link(2, 3)
    return m


def flow_boundary_interpolation_model() -> ModelTables:
    """A cyclic FlowBoundary with block interpolation
    (python/ribasim_testmodels/ribasim_testmodels/basic.py:590-627)."""
    m = _model(end=datetime(2020, 1, 9), interpolation=dict(flow_boundary="block", block_transition_period=0.0))
    m.add_node("FlowBoundary", 1, cyclic_time=True, time=dict(
        time=["2020-01-01 00:00:00", "2020-01-02 00:00:00", "2020-01-03 00:00:00", "2020-01-04 00:00:00"],
        flow_rate=[1e-3, 2.5e-3, 0.0, 1e-3]))
    m.add_node("Basin", 2, state=dict(level=1.0), profile=dict(level=[0.0, 3.0], area=[1000.0, 1000.0]))
    m.add_node("TabulatedRatingCurve", 3, static=dict(level=[0.0, 2.0], flow_rate=[0.0, 1e-3]))
    m.add_node("Terminal", 4)
    for a, b in [(1, 2), (2, 3), (3, 4)]:
        m.add_link(a, b)
    return m


def transient_pump_outlet_model() -> ModelTables:
    """Pump and Outlet with time-dependent flow rates
    (python/ribasim_testmodels/ribasim_testmodels/time.py:59-96)."""
    m = _model()
    time = ["2020-01-01 00:00:00", "2020-07-01 00:00:00", "2021-01-01 00:00:00"]
    flow_rate = [0.0, 1e-5, 1e-5]
    m.add_node("LevelBoundary", 1, static=dict(level=1.1))
    m.add_node("Outlet", 2, time=dict(time=time, flow_rate=flow_rate))
    m.add_node("Basin", 3, state=dict(level=1.0), profile=dict(level=[0.0, 2.0], area=[100.0, 100.0]))
    m.add_node("Pump", 4, time=dict(time=time, flow_rate=flow_rate))
    m.add_node("Terminal", 5)
    for a, b in [(1, 2), (2, 3), (1, 4), (4, 5)]:
        m.add_link(a, b)
    return m


def cyclic_time_model() -> ModelTables:
    """Cyclic forcing, LevelBoundary and FlowBoundary series over a century
    (python/ribasim_testmodels/ribasim_testmodels/basic.py:410-470)."""
    m = _model(end=datetime(2121, 1, 1), solver=dict(saveat=7 * 24 * 60 * 60.0),
               interpolation=dict(flow_boundary="linear"))
    m.add_node("Basin", 1, cyclic_time=True, profile=dict(level=[0.0, 1.0], area=[100.0, 100.0]),
               time=dict(time=["2020-01-01 00:00:00", "2020-04-01 00:00:00", "2020-07-01 00:00:00",
                               "2020-10-01 00:00:00", "2021-01-01 00:00:00"],
                         precipitation=[10.0, 20.0, 10.0, 20.0, 10.0]),
               state=dict(level=5.0))
    m.add_node("LinearResistance", 2, static=dict(resistance=1.0))
    m.add_node("LevelBoundary", 3, cyclic_time=True, time=dict(
        time=["2020-01-01 00:00:00", "2020-05-01 00:00:00", "2020-10-01 00:00:00"], level=[2.0, 3.0, 2.0]))
    m.add_node("FlowBoundary", 4, cyclic_time=True, time=dict(
        time=["2020-01-01 00:00:00", "2020-07-01 00:00:00", "2020-08-01 00:00:00"], flow_rate=[10.0, 20.0, 10.0]))
    for a, b in [(1, 2), (2, 3), (4, 1)]:
        m.add_link(a, b)
    return m


def drought_model() -> ModelTables:
    """A subsection of the LHM Vechtstromen model with a Basin that runs dry (#2189)
    (python/ribasim_testmodels/ribasim_testmodels/basic.py:472-586)."""
    m = _model()
    t0 = "2020-01-01 00:00:00"
    hi = dict(level=[22.4, 22.41, 25.4])
    lo = dict(level=[17.25, 17.26, 20.25])
    m.add_node("Basin", 1558, profile=dict(hi, area=[0.1, 435363.1, 435363.1]), state=dict(level=25.4),
               time=dict(time=[t0], infiltration=[0.0379294629649877049]))
    m.add_node("Basin", 1737, profile=dict(hi, area=[0.1, 422367.9, 422367.9]), state=dict(level=25.4),
               time=dict(time=[t0], infiltration=[0.001642597244767027157]))
    m.add_node("Basin", 2117, profile=dict(level=[17.24, 17.25, 20.24], area=[0.1, 960850.7, 960850.7]),
               state=dict(level=20.24))
    m.add_node("Basin", 2188, profile=dict(lo, area=[0.1, 424545.6, 424545.6]), state=dict(level=20.25))
    m.add_node("Basin", 2189, profile=dict(hi, area=[0.1, 66326.8, 66326.8]), state=dict(level=25.4),
               time=dict(time=[t0], infiltration=[0.0165766882781172575]))
    m.add_node("Basin", 2305, profile=dict(lo, area=[0.1, 9944437.9, 9944437.9]), state=dict(level=20.25))
    for nid, length in ((1236, 3390.0), (1237, 1430.0), (1238, 2710.0)):
        m.add_node("ManningResistance", nid, static=dict(length=length, profile_width=25.0,
                                                         profile_slope=1.0, manning_n=0.04))
    m.add_node("Outlet", 229, static=dict(flow_rate=2.44, min_upstream_level=25.4))
    m.add_node("Outlet", 285, static=dict(flow_rate=5.62, min_upstream_level=20.25))
    for a, b in [(229, 2189), (285, 2117), (1236, 2117), (1237, 2188), (1238, 2188), (1737, 229),
                 (2188, 285), (2305, 1236), (2189, 1237), (1558, 1238)]:
        m.add_link(a, b)
    return m


def circular_flow_model() -> ModelTables:
    """Boezem / polder loop with inlets, a Manning reach and a Pump under hysteresis control
    (python/ribasim_testmodels/ribasim_testmodels/discrete_control.py:801-938)."""
    m = _model(solver=dict(saveat=3600.0))
    days = [m.starttime + timedelta(days=k) for k in range(367)]
    precipitation = np.zeros(367)
    precipitation[0:90] = 1e-6
    precipitation[180:270] = 1e-6
    evaporation = np.zeros(367)
    evaporation[90:180] = 1e-6
    evaporation[270:366] = 1e-6
    for node_id in (3, 4, 6, 9):
        m.add_node("Basin", node_id, profile=dict(area=[10.0, 10_000.0], level=[-10.0, 1.0]),
                   time=dict(time=[d.strftime(TIME_FMT) for d in days], drainage=0.0,
                             potential_evaporation=[float(x) for x in evaporation], infiltration=0.0,
                             precipitation=[float(x) for x in precipitation]),
                   state=dict(level=0.9))
    m.add_node("Outlet", 10, static=dict(flow_rate=10.0, min_upstream_level=1.1))
    m.add_node("Outlet", 12, static=dict(flow_rate=10.0))
    m.add_node("Outlet", 5, static=dict(flow_rate=2.0, min_upstream_level=1.0, max_downstream_level=1.0))
    m.add_node("Outlet", 13, static=dict(flow_rate=2.0, min_upstream_level=1.0, max_downstream_level=0.9))
    m.add_node("ManningResistance", 2, static=dict(length=900.0, manning_n=0.04, profile_width=6.0,
                                                   profile_slope=3.0))
    m.add_node("DiscreteControl", 1,
               variable=dict(listen_node_id=6, variable="level", compound_variable_id=1),
               condition=dict(threshold_high=0.95, threshold_low=0.9, compound_variable_id=1, condition_id=1),
               logic=dict(truth_state=["T", "F"], control_state=["on", "off"]))
    m.add_node("Pump", 7, static=dict(control_state=["on", "off"], flow_rate=[0.1, 0.0]))
    m.add_node("LevelBoundary", 11, static=dict(level=1.1))
    m.add_node("LevelBoundary", 17, static=dict(level=0.9))
    for a, b in [(2, 9), (3, 5), (3, 2), (5, 4), (4, 13), (13, 6), (6, 7), (7, 9), (9, 10), (11, 12),
                 (12, 3), (10, 17), (1, 7)]:
        m.add_link(a, b)
    return m


def tabulated_rating_curve_model() -> ModelTables:
    """One static and one time-varying (cyclic) rating curve between two basins
    (basic.py:276-361)."""
    m = _model()
    prof = dict(area=[0.01, 1000.0], level=[0.0, 1.0])
    m.add_node("Basin", 1, profile=prof, static=dict(precipitation=0.002 / 86400),
               state=dict(level=0.04471158417652035))
    m.add_node("TabulatedRatingCurve", 2, static=dict(level=[0.0, 1.0], flow_rate=[0.0, 10 / 86400]))
    m.add_node("TabulatedRatingCurve", 3, cyclic_time=True, time=dict(
        time=["2020-01-01 00:00:00.000", "2020-01-01 00:00:00.000", "2020-02-01 00:00:00.001",
              "2020-02-01 00:00:00.001", "2020-03-01 00:00:00.000", "2020-03-01 00:00:00.000",
              "2020-04-01 00:00:00.000", "2020-04-01 00:00:00.000"],
        level=[0.0, 1.0, 0.0, 1.1, 0.0, 1.2, 0.0, 1.0], flow_rate=4 * [0.0, 10 / 86400]))
    m.add_node("Basin", 4, profile=prof, static=dict(precipitation=0.0),
               state=dict(level=0.04471158417652035))
    for a, b in ((1, 2), (1, 3), (2, 4), (3, 4)):
        m.add_link(a, b)
    return m


def junction_combined_model() -> ModelTables:
    """Confluence and bifurcation through Junction nodes (junction.py:13-98)."""
    m = _model()
    prof = dict(area=[1000.0, 1000.0], level=[0.0, 1.0])
    m.add_node("Basin", 1, profile=prof, static=dict(infiltration=2.5), state=dict(level=1.0))
    m.add_node("Junction", 2)
    m.add_node("LinearResistance", 3, static=dict(resistance=1.0))
    m.add_node("LinearResistance", 4, static=dict(resistance=1.0))
    m.add_node("Basin", 5, profile=prof, static=dict(surface_runoff=1.0), state=dict(level=1.0))
    m.add_node("Basin", 6, profile=prof, static=dict(drainage=4.0), state=dict(level=1.0))
    m.add_node("LinearResistance", 7, static=dict(resistance=1.0))
    m.add_node("LinearResistance", 8, static=dict(resistance=1.0))
    m.add_node("Junction", 9)
    m.add_node("Basin", 10, profile=prof, static=dict(infiltration=2.5), state=dict(level=1.0))
    for a, b in ((1, 2), (2, 3), (2, 4), (3, 5), (4, 6), (5, 7), (6, 8), (7, 9), (8, 9), (9, 10)):
        m.add_link(a, b)
    return m


def junction_chained_model() -> ModelTables:
    """Basin 1 -> Junction 2 -> Junction 3 -> LR 4/5 -> Basin 6/7 (junction.py:101-165)."""
    m = _model()
    prof = dict(area=[1000.0, 1000.0], level=[0.0, 1.0])
    m.add_node("Basin", 1, profile=prof, static=dict(drainage=1.0), state=dict(level=1.0))
    m.add_node("Junction", 2)
    m.add_node("Junction", 3)
    m.add_node("LinearResistance", 4, static=dict(resistance=1.0))
    m.add_node("LinearResistance", 5, static=dict(resistance=1.0))
    m.add_node("Basin", 6, profile=prof, state=dict(level=0.0))
    m.add_node("Basin", 7, profile=prof, state=dict(level=0.0))
    for a, b in ((1, 2), (2, 3), (3, 4), (3, 5), (4, 6), (5, 7)):
        m.add_link(a, b)
    return m


# ----------------------------------------------------------------------------- synthetic networks
def _random_profile(rng: np.random.Generator, n_points: int, mu: float, sigma: float,
                    bottom: float | None = None):
    level = np.concatenate([[0.0], np.cumsum(rng.uniform(0.5, 2.0, size=n_points - 1))])
    area = np.sort(rng.lognormal(mu, sigma, size=n_points))
    area = np.maximum.accumulate(area + np.arange(n_points) * 1e-3)
    return level + (rng.uniform(-2.0, 2.0) if bottom is None else bottom), area


def _rating_curve(rng: np.random.Generator, bottom: float, n_points: int = 4):
    level = bottom + np.concatenate([[0.0], np.cumsum(rng.uniform(0.3, 1.0, size=n_points - 1))])
    flow = np.concatenate([[0.0], np.cumsum(rng.uniform(0.5, 5.0, size=n_points - 1))])
    return level, flow


def _add_connector(m: ModelTables, rng, kind: str, node_id: int, bottom_src: float) -> None:
    if kind == "ManningResistance":
        m.add_node(kind, node_id, static=dict(
            length=float(rng.uniform(500.0, 5000.0)), manning_n=float(rng.uniform(0.02, 0.05)),
            profile_width=float(rng.uniform(5.0, 50.0)), profile_slope=0.0))
    elif kind == "LinearResistance":
        m.add_node(kind, node_id, static=dict(resistance=float(rng.uniform(50.0, 5000.0)),
                                              max_flow_rate=float(rng.uniform(5.0, 50.0))))
    elif kind == "TabulatedRatingCurve":
        level, flow = _rating_curve(rng, bottom_src)
        m.add_node(kind, node_id, static=dict(level=level, flow_rate=flow))
    elif kind == "Outlet":
        m.add_node(kind, node_id, static=dict(flow_rate=float(rng.uniform(0.5, 10.0)),
                                              min_upstream_level=bottom_src + 0.2))
    elif kind == "Pump":
        m.add_node(kind, node_id, static=dict(flow_rate=float(rng.uniform(0.1, 5.0))))
    else:
        raise ValueError(kind)


def hws_like_model(n_basins: int = 500, seed: int = 20260119, n_days: int = 30,
                   n_pid: int = 10) -> ModelTables:
    """HWS-like synthetic network: river-like tree plus 10 % chords,
    connector mix 35 % Manning / 15 % LR / 20 % TRC / 15 % Outlet / 10 % Pump / 5 % UserDemand,
    20 LevelBoundary, 20 FlowBoundary, `n_pid` PID-controlled outlets, daily forcing."""
    rng = np.random.default_rng(seed)
    m = ModelTables(starttime=datetime(2023, 1, 20),
                    endtime=datetime(2023, 1, 20) + timedelta(days=n_days),
                    solver=dict(algorithm="ImplicitEuler", dt=3600.0))
    days = [m.starttime + timedelta(days=k) for k in range(n_days + 1)]
    # river-like tree: parent close in index, degree bounded; bed levels rise away from the
    # outlet (basin 0) so that the initial state is near a physically meaningful flow regime
    parents = np.zeros(n_basins, dtype=np.int64)
    bottoms = np.zeros(n_basins)
    for b in range(1, n_basins):
        parents[b] = int(rng.integers(max(0, b - 6), b))
        bottoms[b] = bottoms[parents[b]] + rng.uniform(0.02, 0.3)
    nid = 1
    basin_id = []
    basin_rows_time = []
    for b in range(n_basins):
        k = int(rng.integers(2, 9))
        level, area = _random_profile(rng, k, 13.0, 1.5, bottom=float(bottoms[b]))
        m.add_node("Basin", nid, profile=dict(level=level, area=area),
                   state=dict(level=float(level[0] + rng.uniform(0.8, 1.2))))
        P = rng.lognormal(-1.0, 1.0, size=len(days)) * 0.002 / 86400
        E = (1.0 + 0.5 * np.sin(np.arange(len(days)) / 58.0)) * 0.001 / 86400
        D = rng.lognormal(0.0, 0.5, size=len(days)) * 1e-2
        for d, pr, ev, dr in zip(days, P, E, D):
            basin_rows_time.append(dict(node_id=nid, time=d.strftime(TIME_FMT), precipitation=float(pr),
                                        potential_evaporation=float(ev), drainage=float(dr),
                                        infiltration=0.0, surface_runoff=0.0))
        basin_id.append(nid)
        nid += 1
    m.tables["Basin / time"] = basin_rows_time

    kinds = ["ManningResistance", "LinearResistance", "TabulatedRatingCurve", "Outlet", "Pump",
             "UserDemand"]
    probs = np.array([0.35, 0.15, 0.20, 0.15, 0.10, 0.05])
    pairs: set[tuple[int, int]] = set()
    edges: list[tuple[int, int]] = []
    for b in range(1, n_basins):
        parent = int(parents[b])
        edges.append((b, parent))
        pairs.add((min(b, parent), max(b, parent)))
    n_chords = n_basins // 10
    while n_chords > 0:
        a = int(rng.integers(0, n_basins))
        c = int(rng.integers(max(0, a - 20), min(n_basins, a + 20)))
        if a == c or (min(a, c), max(a, c)) in pairs:
            continue
        pairs.add((min(a, c), max(a, c)))
        edges.append((a, c) if bottoms[a] >= bottoms[c] else (c, a))  # downhill
        n_chords -= 1
    outlets_for_pid = []
    for (a, c) in edges:
        kind = str(rng.choice(kinds, p=probs))
        if kind == "UserDemand":
            m.add_node("UserDemand", nid, static=dict(
                demand=float(rng.uniform(0.01, 0.5)), return_factor=float(rng.uniform(0.2, 0.9)),
                min_level=float(bottoms[a] + 0.3), demand_priority=1))
        else:
            _add_connector(m, rng, kind, nid, float(bottoms[a]))
            if kind == "Outlet":
                outlets_for_pid.append((nid, a))
        m.add_link(basin_id[a], nid)
        m.add_link(nid, basin_id[c])
        nid += 1
    # boundaries
    for k in range(20):
        b = int(rng.integers(0, n_basins))
        m.add_node("LevelBoundary", nid, static=dict(level=float(bottoms[b] + rng.uniform(0.5, 1.5))))
        lb = nid
        nid += 1
        kind = "LinearResistance" if k % 2 == 0 else "TabulatedRatingCurve"
        _add_connector(m, rng, kind, nid, float(bottoms[b]))
        m.add_link(basin_id[b], nid)
        m.add_link(nid, lb)
        nid += 1
    for k in range(20):
        b = int(rng.integers(0, n_basins))
        q = rng.lognormal(2.0, 1.0) * rng.lognormal(0.0, 0.2, size=len(days))
        m.add_node("FlowBoundary", nid, time=dict(time=[d.strftime(TIME_FMT) for d in days],
                                                  flow_rate=q))
        m.add_link(nid, basin_id[b])
        nid += 1
    for (outlet_id, a) in outlets_for_pid[:n_pid]:
        m.add_node("PidControl", nid, static=dict(
            listen_node_id=basin_id[a], target=float(bottoms[a] + 1.0), proportional=-0.5,
            integral=-1e-6, derivative=0.0))
        m.add_link(nid, outlet_id)
        nid += 1
    return m


def random_tree_model(n_basins: int = 100_000, seed: int = 20261019, n_hours: int = 24) -> ModelTables:
    """Random-tree network: `parent[i] = rng.integers(max(0,i-50), i)`,
    connector mix 40 % Manning / 20 % TRC / 20 % Outlet / 10 % Pump / 10 % LR, root drains to a
    LevelBoundary through a TabulatedRatingCurve, 1 % of the leaves get FlowBoundaries.

    The rule makes a tree some thousands of basins deep, so the parameters are those of a drainage
    network that has a flow regime: bed levels rise away from the root, every connector is sized
    for the mean inflow that accumulates above it (pump / outlet capacity, resistance, channel
    width, rating curve), initial levels are a metre above the bed."""
    rng = np.random.default_rng(seed)
    m = ModelTables(starttime=datetime(2023, 1, 1),
                    endtime=datetime(2023, 1, 1) + timedelta(hours=n_hours),
                    solver=dict(algorithm="ImplicitEuler", dt=3600.0))
    parent = np.zeros(n_basins, dtype=np.int64)
    for i in range(1, n_basins):
        parent[i] = rng.integers(max(0, i - 50), i)
    is_leaf = np.ones(n_basins, dtype=bool)
    is_leaf[parent[1:]] = False
    kinds = ["ManningResistance", "TabulatedRatingCurve", "Outlet", "Pump", "LinearResistance"]
    probs = np.array([0.4, 0.2, 0.2, 0.1, 0.1])
    rise = rng.uniform(0.0005, 0.004, size=n_basins)
    bottoms = np.zeros(n_basins)
    for i in range(1, n_basins):
        bottoms[i] = bottoms[parent[i]] + rise[i]
    drainage = rng.uniform(0.0, 2e-4, size=n_basins)
    leaves = np.flatnonzero(is_leaf)
    n_fb = max(1, len(leaves) // 100)
    fb_basin = rng.choice(leaves, size=n_fb, replace=False)
    fb_rate = rng.lognormal(-3.0, 1.0, size=n_fb)
    # mean inflow that accumulates above every basin's downstream connector (children have
    # larger indices than their parents)
    acc = drainage.copy()
    np.add.at(acc, fb_basin, fb_rate)
    for i in range(n_basins - 1, 0, -1):
        acc[parent[i]] += acc[i]
    node = m.node
    tabs = m.tables
    prof_rows, state_rows, static_rows = [], [], []
    for b in range(n_basins):
        k = int(rng.integers(2, 5))
        level, area = _random_profile(rng, k, 11.0, 1.0, bottom=float(bottoms[b]))
        bid = b + 1
        node.append(dict(node_id=bid, node_type="Basin", subnetwork_id=None, route_priority=None,
                         cyclic_time=False))
        for lv, ar in zip(level, area):
            prof_rows.append(dict(node_id=bid, level=float(lv), area=float(ar), storage=None))
        state_rows.append(dict(node_id=bid, level=float(level[0] + rng.uniform(0.8, 1.2))))
        static_rows.append(dict(node_id=bid, precipitation=float(rng.uniform(0, 3e-8)),
                                potential_evaporation=float(rng.uniform(0, 1e-8)),
                                drainage=float(drainage[b]), infiltration=0.0,
                                surface_runoff=0.0))
    tabs["Basin / profile"] = prof_rows
    tabs["Basin / state"] = state_rows
    tabs["Basin / static"] = static_rows
    nid = n_basins + 1
    link_id = 1
    choice = rng.choice(len(kinds), size=n_basins, p=probs)
    for i in range(1, n_basins):
        kind = kinds[int(choice[i])]
        _add_connector_fast(m, rng, kind, nid, float(bottoms[i]), float(acc[i]))
        m.link.append(dict(link_id=link_id, from_node_id=i + 1, to_node_id=nid, link_type="flow"))
        m.link.append(dict(link_id=link_id + 1, from_node_id=nid, to_node_id=int(parent[i]) + 1,
                           link_type="flow"))
        link_id += 2
        nid += 1
    # root -> TRC -> LevelBoundary
    _add_connector_fast(m, rng, "TabulatedRatingCurve", nid, float(bottoms[0]), float(acc[0]))
    node.append(dict(node_id=nid + 1, node_type="LevelBoundary", subnetwork_id=None,
                     route_priority=None, cyclic_time=False))
    tabs.setdefault("LevelBoundary / static", []).append(
        dict(node_id=nid + 1, level=float(bottoms[0] - 1.0)))
    m.link.append(dict(link_id=link_id, from_node_id=1, to_node_id=nid, link_type="flow"))
    m.link.append(dict(link_id=link_id + 1, from_node_id=nid, to_node_id=nid + 1,
                       link_type="flow"))
    link_id += 2
    nid += 2
    for b, q in zip(fb_basin, fb_rate):
        node.append(dict(node_id=nid, node_type="FlowBoundary", subnetwork_id=None,
                         route_priority=None, cyclic_time=False))
        tabs.setdefault("FlowBoundary / static", []).append(dict(node_id=nid, flow_rate=float(q)))
        m.link.append(dict(link_id=link_id, from_node_id=nid, to_node_id=int(b) + 1,
                           link_type="flow"))
        link_id += 1
        nid += 1
    return m


def _add_connector_fast(m: ModelTables, rng, kind: str, node_id: int, bottom_src: float,
                        q_mean: float) -> None:
    """Connector sized for the mean flow `q_mean` that arrives at it; appends rows directly (the
    100k generator is loop-heavy)."""
    m.node.append(dict(node_id=node_id, node_type=kind, subnetwork_id=None, route_priority=None,
                       cyclic_time=False))
    rows = m.tables.setdefault(f"{kind} / static", [])
    q = max(q_mean, 1e-4)
    if kind == "ManningResistance":
        # ~0.3 m3/s per metre of width at 1 m depth and a 0.2 m head over 1 km (n = 0.04)
        rows.append(dict(node_id=node_id, length=float(rng.uniform(200.0, 2000.0)),
                         manning_n=float(rng.uniform(0.02, 0.06)),
                         profile_width=float(min(max(q / 0.3 * rng.uniform(0.7, 1.5), 1.0), 2000.0)),
                         profile_slope=0.0))
    elif kind == "LinearResistance":
        rows.append(dict(node_id=node_id, resistance=float(rng.uniform(0.1, 0.5) / q),
                         max_flow_rate=float(5.0 * q)))
    elif kind == "TabulatedRatingCurve":
        level, flow = _rating_curve(rng, bottom_src)
        flow = flow / flow[2] * q * rng.uniform(0.8, 1.6)  # carries the mean flow at ~1 m depth
        for lv, qq in zip(level, flow):
            rows.append(dict(node_id=node_id, level=float(lv), flow_rate=float(qq)))
    elif kind == "Outlet":
        rows.append(dict(node_id=node_id, flow_rate=float(q * rng.uniform(1.2, 2.5)),
                         min_upstream_level=bottom_src + 0.2))
    elif kind == "Pump":
        rows.append(dict(node_id=node_id, flow_rate=float(q * rng.uniform(1.2, 2.5))))


MODELS = dict(
    tabulated_rating_curve=tabulated_rating_curve_model,
    junction_combined=junction_combined_model,
    junction_chained=junction_chained_model,
    basic=basic_model,
    basic_transient=basic_transient_model,
    backwater=backwater_model,
    trivial=trivial_model,
    bucket=bucket_model,
    leaky_bucket=leaky_bucket_model,
    linear_resistance=linear_resistance_model,
    rating_curve=rating_curve_model,
    rating_curve_between_basins=rating_curve_between_basins_model,
    manning_resistance=manning_resistance_model,
    misc_nodes=misc_nodes_model,
    pid_control_equation=pid_control_equation_model,
    pid_control=pid_control_model,
    outlet_continuous_control=outlet_continuous_control_model,
    user_demand=user_demand_model,
    pump_discrete_control=pump_discrete_control_model,
    flow_boundary_time=flow_boundary_time_model,
    two_basin=two_basin_model,
    discrete_control_of_pid_control=discrete_control_of_pid_control_model,
    manning_discrete_control=manning_discrete_control_model,
    flow_condition=flow_condition_model,
    level_boundary_condition=level_boundary_condition_model,
    tabulated_rating_curve_control=tabulated_rating_curve_control_model,
    compound_variable_condition=compound_variable_condition_model,
    connector_node_flow_condition=connector_node_flow_condition_model,
    storage_condition=storage_condition_model,
    transient_condition=transient_condition_model,
    outlet=outlet_model,
    flow_boundary_interpolation=flow_boundary_interpolation_model,
    transient_pump_outlet=transient_pump_outlet_model,
    drought=drought_model,
    circular_flow=circular_flow_model,
    minimal_subnetwork=minimal_subnetwork_model,
)
