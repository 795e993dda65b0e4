"""BMI surface of the reference (`core/src/bmi.jl:6-99`, exported by libribasim:
`build/libribasim.jl:74-212`) on top of the device-resident `Model`.

Same names, argument meaning and error behaviour as the reference's BMI functions, so that tests
read like `core/test/bmi_test.jl` and `python/ribasim_api/tests/test_bmi.py`.  In the reference
`get_value_ptr` hands out the solver's own arrays; here they are host mirrors of device memory:
arrays the caller may write (`basin.infiltration`, `basin.drainage`, `basin.surface_runoff`,
`user_demand.demand`) are uploaded at the start of `update*`, result arrays are refreshed in place
at its end — the array objects never change, like the reference's pointers.
"""

from __future__ import annotations

from pathlib import Path

import numpy as np

from .model import Model

_WRITABLE = {"basin.infiltration": "infiltration", "basin.drainage": "drainage",
             "basin.surface_runoff": "surface_runoff"}


class RibasimBmi:
    """`BMI.initialize(Ribasim.Model, toml)` ... `BMI.finalize(model)` for one model instance
    (ensemble member 0 of a one-member batch unless `n_members` is given)."""

    def __init__(self) -> None:
        self.model: Model | None = None
        self._values: dict[str, np.ndarray] = {}

    # -- libribasim: initialize / finalize (build/libribasim.jl:74-88)
    def initialize(self, config_path: str | Path, n_members: int = 1, **model_kw) -> None:
        self.model = Model(config_path, n_members=n_members, **model_kw)
        m = self.model
        f = m.flat.ints
        nb, nu = f["n_basin"], f["n_ud_link"]
        # `user_demand.demand` is `vec(user_demand.demand)` of a column-major [node, priority]
        # Matrix (core/src/bmi.jl:82-83, docs/dev/bmi.qmd): node index fastest.  The host mirror is
        # kept in Fortran order so that the flat array handed out is a view with that layout.
        ud = m.p.nodes["user_demand"]
        ud["demand"] = np.asfortranarray(ud["demand"], dtype=np.float64)
        self._values = {
            "basin.storage": np.zeros(nb), "basin.level": np.zeros(nb),
            "basin.cumulative_infiltration": np.zeros(nb),
            "basin.cumulative_drainage": np.zeros(nb),
            "basin.cumulative_surface_runoff": np.zeros(nb),
            "basin.subgrid_level": np.full(len(m.subgrid), np.nan),
            "user_demand.cumulative_inflow": np.zeros(nu),
            "user_demand.demand": ud["demand"].reshape(-1, order="F"),
        }
        for name, key in _WRITABLE.items():
            self._values[name] = m.p.basin.vertical_flux[key]
        self._refresh_outputs()

    def finalize(self) -> None:
        self._require().finalize()
        self.model = None

    def _require(self) -> Model:
        if self.model is None:
            raise RuntimeError("Model not initialized")  # build/libribasim.jl:21-24
        return self.model

    # -- time (core/src/bmi.jl:91-99)
    def get_current_time(self) -> float:
        return self._require().t

    def get_start_time(self) -> float:
        self._require()
        return 0.0

    def get_end_time(self) -> float:
        return self._require().t_end

    def get_time_step(self) -> float:
        m = self._require()
        return m.qndf.h if m.adaptive and m.qndf.h > 0 else m.dt

    def get_time_units(self) -> str:
        return "s"

    # -- stepping (core/src/bmi.jl:36-52)
    def update(self) -> None:
        self._require().update()
        self._refresh_outputs()

    def update_until(self, time: float) -> None:
        m = self._require()
        dt = time - m.t
        if dt < 0:
            raise RuntimeError("The model has already passed the given timestamp.")
        if dt == 0:
            return
        m.update_until(time)
        self._refresh_outputs()

    def update_subgrid_level(self) -> None:
        """libribasim `update_subgrid_level` (build/libribasim.jl:98-101, core/src/bmi.jl:54-57)."""
        m = self._require()
        self._values["basin.subgrid_level"][:] = m.subgrid_level()[:, 0]

    # -- variables (core/src/bmi.jl:59-89; libribasim get_var_type/rank/shape: :115-160)
    def get_value_ptr(self, name: str) -> np.ndarray:
        self._require()
        if name not in self._values:
            raise KeyError(f"Unknown variable {name}")
        return self._values[name]

    def get_var_type(self, name: str) -> str:
        self.get_value_ptr(name)
        return "double"

    def get_var_rank(self, name: str) -> int:
        return self.get_value_ptr(name).ndim

    def get_var_shape(self, name: str) -> tuple[int, ...]:
        return self.get_value_ptr(name).shape

    def get_component_name(self) -> str:
        return "Ribasim"

    def _refresh_outputs(self) -> None:
        m = self._require()
        f = m.flat.ints
        nb = f["n_basin"]
        u = m.h.get_member("u", f["n_state"])[:, 0]
        v = self._values
        v["basin.storage"][:] = m.h.get_member("storage", nb)[:, 0]
        v["basin.level"][:] = m.h.get_member("level", nb)[:, 0]
        v["basin.cumulative_infiltration"][:] = u[f["off_infil"]:f["off_infil"] + nb]
        v["basin.cumulative_drainage"][:] = m.h.get_member("cum_drainage", nb)[:, 0]
        v["basin.cumulative_surface_runoff"][:] = m.h.get_member("cum_surface_runoff", nb)[:, 0]
        v["user_demand.cumulative_inflow"][:] = u[f["off_ud_in"]:f["off_ud_in"] + f["n_ud_link"]]
