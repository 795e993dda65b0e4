"""BASELINE.json configs 1 and 2 run end to end on the GPU (Model driver -> C ABI -> CUDA path)
against the reference's own golden numbers, not only against the oracle."""

from __future__ import annotations

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _model(name, **kw):
    from ribasim_b200.model import Model
    from ribasim_b200.network import build_params
    from ribasim_b200.testmodels import MODELS

    return Model(build_params(MODELS[name]()), dt=3600.0, **kw)


def test_basic_end_of_year_storages():
    # core/test/run_models_test.jl:221-222 (config 1: the reference's correctness case)
    model = _model("basic", n_members=2)
    try:
        model.run()
        S = model.storage()
        for m in range(2):
            assert np.allclose(S[:, m], [775.23576, 775.23365, 572.60102, 1130.005], atol=1.5, rtol=0)
        assert model.status_counts == dict(newton_failed=0, negative_storage=0)
    finally:
        model.finalize()


def test_trivial_end_state():
    # core/regression_test/regression_test.jl:35-41 (ImplicitEuler row)
    model = _model("trivial", n_members=1)
    try:
        model.run()
        assert model.storage()[0, 0] == pytest.approx(62.2290641115, abs=0.02)
        assert model.level()[0, 0] == pytest.approx(0.352778, abs=1e-4)
    finally:
        model.finalize()


def test_backwater_single_instance_against_standard_step_method():
    # core/test/run_models_test.jl:489-523 (config 2: 50 Manning reaches, single instance):
    # levels within 0.02 m of the standard-step backwater curve, flow in = flow out = 5 +- 1e-3
    from tests.test_oracle_golden import _standard_step_method

    model = _model("backwater", n_members=1)
    try:
        model.run()
        flat = model.flat
        lv = model.level()[:, 0]
        h_actual = lv[:50]
        x = np.arange(10.0, 991.0, 20.0)
        h_expected = _standard_step_method(x, 5.0, 1.0, 0.04, h_actual[-1], 1.0e-6)
        assert np.all(np.abs(h_expected - h_actual) < 0.02)
        model.h.refresh(model.t)
        du = model.flow()[:, 0]
        assert du[flat["off_manning"] + 49] == pytest.approx(5.0, abs=1e-3)
        assert du[flat["off_manning"]] == pytest.approx(5.0, abs=1e-3)
        assert model.status_counts == dict(newton_failed=0, negative_storage=0)
    finally:
        model.finalize()
