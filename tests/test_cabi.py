"""The drop-in boundary: librbs.so loads and exports every symbol include/ribasim_b200.h
declares; without a CUDA device it fails loudly instead of falling back to a CPU path."""

from __future__ import annotations

import ctypes
import re
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parents[1]


def _declared_symbols() -> list[str]:
    text = (ROOT / "include" / "ribasim_b200.h").read_text()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(rbs_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_are_exported():
    from ribasim_b200 import lib

    L = lib.load()
    declared = _declared_symbols()
    assert len(declared) >= 26
    for name in declared:
        assert hasattr(L, name), f"{name} declared in include/ribasim_b200.h but not exported"
    assert sorted(lib.EXPORTED_SYMBOLS) == declared
    assert b"sm_100a" in L.rbs_version()


def test_library_is_sm100a_only():
    import subprocess

    from ribasim_b200 import lib

    out = subprocess.run(["cuobjdump", "--list-elf", str(lib.LIB_PATH)], capture_output=True,
                         text=True)
    if out.returncode != 0:
        pytest.skip("cuobjdump unavailable")
    archs = set(re.findall(r"sm_\d+a?", out.stdout))
    assert archs == {"sm_100a"}, archs


def test_builder_argument_errors():
    from ribasim_b200 import lib

    L = lib.load()
    b = L.rbs_builder_create()
    assert L.rbs_builder_set_i32(b, b"n_basin", None, 3) != 0
    assert "null" in lib.last_error()
    L.rbs_builder_destroy(b)
    assert not L.rbs_create(None, 1, 0)
    assert "null builder" in lib.last_error()


def test_no_cpu_fallback_without_device():
    import torch

    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    from ribasim_b200 import lib
    from ribasim_b200.flatten import flatten
    from ribasim_b200.network import build_params
    from ribasim_b200.testmodels import MODELS

    flat = flatten(build_params(MODELS["basic"]()))
    with pytest.raises(lib.RbsError, match="no usable CUDA device"):
        lib.Handle(flat, n_members=2)


def test_product_package_does_not_import_oracle():
    """Only tests/, smoke() and bench.py's CPU-baseline legs may touch oracle/."""
    for path in (ROOT / "ribasim_b200").rglob("*"):
        if path.suffix in (".py", ".cu", ".cuh", ".cpp", ".h"):
            text = path.read_text()
            assert "oracle" not in re.sub(r"#.*|//.*", "", text).replace("CPU oracle", ""), path
