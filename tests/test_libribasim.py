"""The `libribasim` stand-in (ribasim_b200/libribasim.so) bound ONLY by the names and signatures of
the reference's C ABI (/root/reference/build/libribasim.jl:74-212, include/libribasim.h), the way
xmipy / python/ribasim_api binds the reference library, replaying
/root/reference/python/ribasim_api/tests/test_bmi.py.  `-m "not gpu"`: exports, error contract and a
plain C host without a model; `-m gpu`: the BMI walk-through on the device.
"""

from __future__ import annotations

import ctypes
import re
import subprocess
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent

SYMBOLS = ["initialize", "finalize", "update", "update_until", "update_subgrid_level",
           "get_current_time", "get_start_time", "get_end_time", "get_time_step", "get_var_type",
           "get_var_rank", "get_value_ptr", "get_value_ptr_double", "get_var_shape",
           "get_component_name", "get_version", "get_last_bmi_error", "execute",
           "ribasim_set_julia_threads"]


@pytest.fixture(scope="module")
def shim_path():
    from ribasim_b200 import lib

    return lib.build_shim()


class XmiError(RuntimeError):
    pass


class Xmi:
    """What xmipy.XmiWrapper + ribasim_api.RibasimApi do with the library
    (python/ribasim_api/ribasim_api/ribasim_api.py:5-26): ctypes.CDLL, byte buffers of the BMI_LEN*
    sizes, nonzero return code -> exception carrying get_last_bmi_error."""

    LENVARTYPE, LENCOMPONENTNAME, LENVERSION, LENERRMESSAGE = 51, 256, 256, 1025

    def __init__(self, path):
        self.lib = ctypes.CDLL(str(path))

    def _check(self, rc, what):
        if rc != 0:
            buf = ctypes.create_string_buffer(self.LENERRMESSAGE)
            self.lib.get_last_bmi_error(buf)
            raise XmiError(f"BMI exception in {what}: Message from Ribasim '{buf.value.decode()}'")

    def initialize(self, config_file):
        self._check(self.lib.initialize(ctypes.c_char_p(str(config_file).encode())), "initialize")

    def finalize(self):
        self._check(self.lib.finalize(), "finalize")

    def update(self):
        self._check(self.lib.update(), "update")

    def update_until(self, time):
        self._check(self.lib.update_until(ctypes.c_double(time)), "update_until")

    def update_subgrid_level(self):
        self._check(self.lib.update_subgrid_level(), "update_subgrid_level")

    def execute(self, config_file):
        self._check(self.lib.execute(ctypes.c_char_p(str(config_file).encode())), "execute")

    def _time(self, fn):
        v = ctypes.c_double(0.0)
        self._check(getattr(self.lib, fn)(ctypes.byref(v)), fn)
        return v.value

    def get_start_time(self):
        return self._time("get_start_time")

    def get_current_time(self):
        return self._time("get_current_time")

    def get_end_time(self):
        return self._time("get_end_time")

    def get_time_step(self):
        return self._time("get_time_step")

    def get_var_type(self, name):
        buf = ctypes.create_string_buffer(self.LENVARTYPE)
        self._check(self.lib.get_var_type(ctypes.c_char_p(name.encode()), buf), "get_var_type")
        return buf.value.decode()

    def get_var_rank(self, name):
        r = ctypes.c_int(0)
        self._check(self.lib.get_var_rank(ctypes.c_char_p(name.encode()), ctypes.byref(r)), "get_var_rank")
        return r.value

    def get_var_shape(self, name):
        rank = self.get_var_rank(name)
        arr = np.zeros(rank, dtype=np.int32)
        self._check(self.lib.get_var_shape(ctypes.c_char_p(name.encode()),
                                           arr.ctypes.data_as(ctypes.POINTER(ctypes.c_int))), "get_var_shape")
        return arr

    def get_value_ptr(self, name):
        shape = self.get_var_shape(name)
        assert self.get_var_type(name) == "double"
        ptr = ctypes.POINTER(ctypes.c_double)()
        self._check(self.lib.get_value_ptr(ctypes.c_char_p(name.encode()), ctypes.byref(ptr)), "get_value_ptr")
        if int(np.prod(shape)) == 0:
            return np.zeros(tuple(shape))
        return np.ctypeslib.as_array(ptr, shape=tuple(int(x) for x in shape))

    def get_component_name(self):
        buf = ctypes.create_string_buffer(self.LENCOMPONENTNAME)
        self._check(self.lib.get_component_name(buf), "get_component_name")
        return buf.value.decode()

    def get_version(self):
        buf = ctypes.create_string_buffer(self.LENVERSION)
        self._check(self.lib.get_version(buf), "get_version")
        return buf.value.decode()


# ----------------------------------------------------------------------------- CPU
def test_header_and_exports_match(shim_path):
    """Every symbol of the reference library is declared in include/libribasim.h and exported."""
    header = (ROOT / "include" / "libribasim.h").read_text()
    declared = re.findall(r"^int (\w+)\(", header, flags=re.M)
    assert sorted(declared) == sorted(SYMBOLS)
    out = subprocess.run(["nm", "-D", "--defined-only", str(shim_path)], capture_output=True, text=True, check=True)
    exported = sorted(line.split()[-1] for line in out.stdout.splitlines() if " T " in line)
    assert exported == sorted(SYMBOLS)


def test_try_c_contract_without_model(shim_path):
    """@try_c (libribasim.jl:40-52): 1 + message before initialize; the *_uninitialized calls work."""
    x = Xmi(shim_path)
    x.finalize()  # fine without a model (libribasim.jl:82-90)
    assert x.get_component_name() == "Ribasim"
    assert x.get_version() == "2026.1.2"
    for call in (x.update, lambda: x.update_until(1.0), x.get_current_time, x.get_end_time,
                 lambda: x.get_var_type("basin.storage"), x.update_subgrid_level):
        with pytest.raises(XmiError, match="Model not initialized"):
            call()
    with pytest.raises(XmiError):
        x.initialize("/nonexistent/ribasim.toml")
    with pytest.raises(XmiError, match="Model not initialized"):
        x.update()
    assert x.lib.ribasim_set_julia_threads(ctypes.c_int32(2)) == 1  # the runtime is already up


def test_plain_c_host_without_model(shim_path, tmp_path):
    """A process without Python binds the library with dlopen like build/cli/src/main.rs:59-87."""
    exe = tmp_path / "capi_host"
    subprocess.run(["gcc", "-O1", "-o", str(exe), str(ROOT / "tests" / "native" / "capi_host.c"), "-ldl"], check=True)
    out = subprocess.run([str(exe), str(shim_path)], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stdout + out.stderr
    lines = dict(line.split(" ", 1) for line in out.stdout.strip().splitlines())
    assert lines["component"] == "Ribasim" and lines["version"] == "2026.1.2"
    assert lines["set_threads_range"] == "2" and lines["set_threads"] == "0" and lines["set_threads_late"] == "1"
    assert lines["update_uninitialized"] == "1 Model not initialized"
    assert lines["finalize_uninitialized"] == "0"


# ----------------------------------------------------------------------------- GPU
def _write(name, tmp_path, **solver):
    from ribasim_b200.testmodels import MODELS

    m = MODELS[name]()
    m.solver.update(algorithm="ImplicitEuler", dt=60.0, **solver)
    return m.write(tmp_path / name)


@pytest.fixture
def api(shim_path):
    x = Xmi(shim_path)
    yield x
    x.finalize()


@pytest.mark.gpu
def test_bmi_times_and_update(api, tmp_path):
    # test_bmi.py:11-57
    api.initialize(_write("basic", tmp_path))
    assert api.get_start_time() == pytest.approx(0.0)
    assert api.get_current_time() == pytest.approx(api.get_start_time())
    assert api.get_end_time() == pytest.approx(366 * 86400.0)
    api.update()
    assert api.get_current_time() > 0.0
    api.update_until(3600.0)
    assert api.get_current_time() == pytest.approx(3600.0)
    with pytest.raises(XmiError, match="already passed"):
        api.update_until(60.0)


@pytest.mark.gpu
def test_bmi_variables(api, tmp_path):
    # test_bmi.py:72-110, 185-199
    api.initialize(_write("basic", tmp_path))
    assert api.get_var_type("basin.storage") == "double"
    assert api.get_var_rank("basin.storage") == 1
    assert np.array_equal(api.get_var_shape("basin.storage"), [4])
    np.testing.assert_array_almost_equal(api.get_value_ptr("basin.storage"), np.ones(4))
    np.testing.assert_array_almost_equal(api.get_value_ptr("basin.level"), 4 * [0.04471158417652035])
    name = "unknown_node.unknown_variable"
    with pytest.raises(XmiError, match=re.escape(f"Message from Ribasim 'Unknown variable {name}'")):
        api.get_var_type(name)
    # the pointer is the model's own array: stable, refreshed in place by update
    s = api.get_value_ptr("basin.storage")
    before = s.copy()
    api.update_until(86400.0)
    assert api.get_value_ptr("basin.storage").ctypes.data == s.ctypes.data
    assert not np.array_equal(s, before)


@pytest.mark.gpu
def test_bmi_subgrid(api, tmp_path):
    # test_bmi.py:60-70, 113-126
    api.initialize(_write("two_basin", tmp_path))
    api.update_until(86400.0)
    api.update_subgrid_level()
    level = api.get_value_ptr("basin.subgrid_level")
    assert np.isfinite(level).all()
    np.testing.assert_allclose(level, api.get_value_ptr("basin.level"), rtol=1e-12)


@pytest.mark.gpu
def test_bmi_basin_forcing_and_writable_pointer(api, tmp_path):
    # test_bmi.py:129-156 and core/test/bmi_test.jl (writable infiltration / drainage)
    api.initialize(_write("leaky_bucket", tmp_path))
    np.testing.assert_array_almost_equal(api.get_value_ptr("basin.infiltration"), [0.0])
    drainage = api.get_value_ptr("basin.drainage")
    np.testing.assert_array_almost_equal(drainage, [0.003])
    api.update_until(60.0)
    np.testing.assert_array_almost_equal(api.get_value_ptr("basin.cumulative_infiltration"), [0.0])
    np.testing.assert_array_almost_equal(api.get_value_ptr("basin.cumulative_drainage"), [0.003 * 60.0])
    drainage[0] = 0.006  # a coupler writes through the pointer: the next steps use it
    api.update_until(120.0)
    np.testing.assert_array_almost_equal(api.get_value_ptr("basin.cumulative_drainage"), [0.003 * 60.0 + 0.006 * 60.0])


@pytest.mark.gpu
def test_bmi_user_demand(api, tmp_path):
    # test_bmi.py:159-174
    api.initialize(_write("user_demand", tmp_path))
    np.testing.assert_array_almost_equal(api.get_value_ptr("user_demand.demand"), [1e-4, 0.0, 0.0])
    api.update_until(60.0)
    np.testing.assert_array_almost_equal(api.get_value_ptr("user_demand.cumulative_inflow"), [1e-4 * 60.0, 0.0, 0.0])


@pytest.mark.gpu
def test_execute_writes_results(shim_path, tmp_path):
    # test_bmi.py:205-208
    from ribasim_b200.testmodels import MODELS

    m = MODELS["basic"]()
    m.solver.update(algorithm="ImplicitEuler", dt=86400.0)
    m.endtime = m.starttime.replace(month=2)
    toml = m.write(tmp_path / "basic")
    Xmi(shim_path).execute(toml)
    assert (Path(toml).parent / "results" / "basin.nc").exists()


@pytest.mark.gpu
def test_plain_c_host_runs_a_model(shim_path, tmp_path):
    exe = tmp_path / "capi_host"
    subprocess.run(["gcc", "-O1", "-o", str(exe), str(ROOT / "tests" / "native" / "capi_host.c"), "-ldl"], check=True)
    toml = _write("basic", tmp_path)
    out = subprocess.run([str(exe), str(shim_path), str(toml)], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stdout + out.stderr
    lines = dict(line.split(" ", 1) for line in out.stdout.strip().splitlines())
    assert lines["initialize"] == "0" and lines["bad"] == "0"
    assert lines["var_type"] == "double" and lines["rank_shape"] == "1 4"
    assert np.allclose([float(v) for v in lines["storage0"].split()], 1.0)
    assert float(lines["after_update"]) == 60.0 and float(lines["after_update_until"]) == 86400.0
    assert lines["pointer_stable"] == "1"
    assert lines["unknown_var"] == "1 Unknown variable unknown_node.unknown_variable"
    assert lines["backwards"].startswith("1 ") and "already passed" in lines["backwards"]
    assert lines["update_finalized"] == "1 Model not initialized"


@pytest.mark.gpu
def test_default_toml_runs_adaptive_qndf(api, tmp_path):
    """The reference's test models ask for its default algorithm, QNDF (python/ribasim_api/tests/
    test_bmi.py writes them unchanged): `update` takes one adaptive step, `update_until` lands
    exactly on the requested time (test_bmi.py:41-57)."""
    from ribasim_b200.testmodels import MODELS

    toml = MODELS["basic"]().write(tmp_path / "basic_default")
    api.initialize(toml)
    api.update()
    t1 = api.get_current_time()
    assert 0.0 < t1 < 3600.0           # the first adaptive step is short
    assert api.get_time_step() > 0.0
    api.update_until(60.0)
    assert api.get_current_time() == pytest.approx(60.0)
    api.update_until(86400.0)
    assert api.get_current_time() == pytest.approx(86400.0)
    s = api.get_value_ptr("basin.storage")
    assert np.all(np.isfinite(s)) and np.all(s > 0)
