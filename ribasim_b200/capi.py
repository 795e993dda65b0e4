"""Python side of the `libribasim` stand-in (`csrc/libribasim_shim.c`).

The shim exports the 19 C symbols of the reference's shared library
(`/root/reference/build/libribasim.jl:74-212`) and forwards each of them to one function of this
module, which keeps the reference's global `model` (`libribasim.jl:5-7`) as a `RibasimBmi` on top
of the device-resident `Model`.  Functions return plain Python values; exceptions propagate to the
shim, which turns them into return code 1 and the `get_last_bmi_error` message (`@try_c`,
`libribasim.jl:40-52`).  The arrays handed out by `value_address` are the BMI host mirrors of
`bmi.py`: they live as long as the model does, like the reference's pointers.
"""

from __future__ import annotations

import numpy as np

from .bmi import RibasimBmi

RIBASIM_VERSION = "2026.1.2"  # the reference version this path mirrors (core/Project.toml)

_bmi: RibasimBmi | None = None


def _model() -> RibasimBmi:
    if _bmi is None or _bmi.model is None:
        raise RuntimeError("Model not initialized")  # libribasim.jl:21-24
    return _bmi


def initialize(path: str) -> int:
    """`initialize(path::Cstring)` (libribasim.jl:74-80): a new global model replaces the old."""
    global _bmi
    if _bmi is not None and _bmi.model is not None:
        _bmi.finalize()
    _bmi = None
    b = RibasimBmi()
    b.initialize(path)
    _bmi = b
    return 0


def finalize() -> int:
    """`finalize()` (libribasim.jl:82-90): fine without a model."""
    global _bmi
    if _bmi is not None and _bmi.model is not None:
        _bmi.finalize()
    _bmi = None
    return 0


def _retcode(b: RibasimBmi, failed_before: int) -> int:
    # `update_retcode` (libribasim.jl:9-11): nonzero when the integrator did not succeed
    return int(b.model.status_counts["newton_failed"] > failed_before)


def update() -> int:
    b = _model()
    f0 = b.model.status_counts["newton_failed"]
    b.update()
    return _retcode(b, f0)


def update_until(time: float) -> int:
    b = _model()
    f0 = b.model.status_counts["newton_failed"]
    b.update_until(float(time))
    return _retcode(b, f0)


def update_subgrid_level() -> int:
    _model().update_subgrid_level()
    return 0


def get_current_time() -> float:
    return float(_model().get_current_time())


def get_start_time() -> float:
    return float(_model().get_start_time())


def get_end_time() -> float:
    return float(_model().get_end_time())


def get_time_step() -> float:
    return float(_model().get_time_step())


def _value(name: str) -> np.ndarray:
    b = _model()
    try:
        return b.get_value_ptr(name)
    except KeyError:
        raise RuntimeError(f"Unknown variable {name}") from None  # core/src/bmi.jl:86


def get_var_type(name: str) -> str:
    _value(name)
    return "double"  # c_type_name(Float64), libribasim.jl:216-218


def get_var_rank(name: str) -> int:
    return int(_value(name).ndim)


def get_var_shape(name: str) -> tuple[int, ...]:
    return tuple(int(x) for x in _value(name).shape)


def value_address(name: str) -> int:
    """Address of the first element of the variable's host array (`pointer(value)`,
    libribasim.jl:142-149)."""
    v = _value(name)
    assert v.dtype == np.float64
    return int(v.ctypes.data)


def get_component_name() -> str:
    return "Ribasim"


def get_version() -> str:
    return RIBASIM_VERSION


def execute(path: str) -> int:
    """`execute(toml_path)` = `Ribasim.main` (libribasim.jl:190-192, core/src/main.jl:10-69): run
    the model to its end time and write the result files; 0 on success."""
    from pathlib import Path

    from .model import Model

    model = Model(path)
    try:
        results_dir = Path(path).resolve().parent / str(model.tables.results_dir)
        model.run(results_dir=results_dir)
        ok = model.status_counts["newton_failed"] == 0
    finally:
        model.finalize()
    return 0 if ok else 1
