"""CPU-side checks: the C-ABI library loads and exports every symbol the header declares; the
host-side mirror of the reference interface validates like the reference; the product never
imports the oracle; without a GPU the compute entry points fail loudly (no CPU fallback)."""
import ctypes as C
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "varpro_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(vp_[a-z_0-9]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    from varpro_b200 import _lib
    lib = _lib.load()
    declared = _declared_symbols()
    assert len(declared) >= 20
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in include/varpro_b200.h but not exported"
    assert set(declared) == set(_lib.SYMBOLS), "ctypes table and header disagree"
    assert lib.vp_abi_version() == 2
    assert lib.vp_status_string(3).decode() == "x or y must have nonzero number of elements"


def test_struct_layouts_match_header():
    from varpro_b200 import _lib
    assert C.sizeof(_lib.BasisDesc) == 4 + 4 + 16 + 8
    assert C.sizeof(_lib.LmOptions) == 4 * 8 + 8
    assert C.sizeof(_lib.FitReport) == 24
    assert C.sizeof(_lib.Reduced) == 8 + 8 * 8 + 64 * 8 + 8


def test_product_never_touches_the_oracle():
    for dirpath, _, files in os.walk(os.path.join(ROOT, "varpro_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in src.replace("the oracle", "").lower(), f"{f} mentions the oracle directory"


def test_no_cpu_fallback_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    import varpro_b200 as vb
    m = (vb.SeparableModelBuilder(["tau"]).function(["tau"], vb.ExpDecay()).independent_variable(np.linspace(0, 1, 8))
         .initial_parameters([1.0]).build())
    with pytest.raises(vb.VarproError) as ei:
        vb.SeparableProblemBuilder.new(m).observations(np.ones(8)).build()
    assert ei.value.status == 41 and "no CPU fallback" in str(ei.value)


def test_model_builder_errors_mirror_reference():
    """src/model/builder/mod.rs + error.rs: errors are reported by build()."""
    import varpro_b200 as vb
    from varpro_b200 import api
    x = np.linspace(0, 1, 4)
    ok = lambda: vb.SeparableModelBuilder(["a", "b"]).independent_variable(x).initial_parameters([1.0, 2.0])
    with pytest.raises(api.EmptyModel):
        ok().build()
    with pytest.raises(api.UnusedParameter):
        ok().function(["a"], vb.ExpDecay()).build()
    with pytest.raises(api.FunctionParameterNotInModel):
        ok().function(["c"], vb.ExpDecay()).function(["b"], vb.ExpDecay()).build()
    with pytest.raises(api.IncorrectParameterCount):
        ok().function(["a", "b"], vb.ExpDecay()).build()
    with pytest.raises(api.DuplicateParameterNames):
        vb.SeparableModelBuilder(["a", "a"]).function(["a"], vb.ExpDecay()).independent_variable(x).initial_parameters([1, 2]).build()
    with pytest.raises(api.CommaInParameterNameNotAllowed):
        vb.SeparableModelBuilder(["a,b"]).function(["a,b"], vb.ExpDecay()).independent_variable(x).initial_parameters([1]).build()
    with pytest.raises(api.MissingX):
        vb.SeparableModelBuilder(["a"]).function(["a"], vb.ExpDecay()).initial_parameters([1.0]).build()
    with pytest.raises(api.MissingInitialParameters):
        vb.SeparableModelBuilder(["a"]).function(["a"], vb.ExpDecay()).independent_variable(x).build()
    with pytest.raises(api.IncorrectParameterCount):
        vb.SeparableModelBuilder(["a"]).function(["a"], vb.ExpDecay()).independent_variable(x).initial_parameters([1.0, 2.0]).build()
    m = ok().function(["b"], vb.ExpDecay()).function(["a"], vb.ExpDecay()).invariant_function(vb.Constant()).build()
    assert m.parameter_count() == 2 and m.base_function_count() == 3 and m.output_len() == 4
    assert [list(idx) for _, idx in m._functions] == [[1], [0], []]  # create_index_mapping
    with pytest.raises(vb.ModelError):
        m.set_params([1.0])


def test_problem_builder_errors_mirror_reference():
    """src/problem/builder.rs:278-302 and src/problem/builder/test.rs:111-183."""
    import varpro_b200 as vb
    from varpro_b200 import api
    x = np.linspace(0, 1, 4)
    m = vb.SeparableModelBuilder(["a"]).function(["a"], vb.ExpDecay()).independent_variable(x).initial_parameters([1.0]).build()
    with pytest.raises(api.YDataMissing):
        vb.SeparableProblemBuilder.new(m).build()
    with pytest.raises(api.InvalidLengthOfData):
        vb.SeparableProblemBuilder.new(m).observations(np.ones(5)).build()
    with pytest.raises(api.ZeroLengthVector):
        vb.SeparableProblemBuilder.new(m).observations(np.ones(0)).build()
    with pytest.raises(api.InvalidLengthOfWeights):
        vb.SeparableProblemBuilder.new(m).observations(np.ones(4)).weights(np.ones(3)).build()
    b = vb.SeparableProblemBuilder.new(m).epsilon(-3.0)
    assert b._eps == 3.0  # epsilon = |epsilon| (builder.rs:248)


def test_levenberg_marquardt_option_holder():
    import varpro_b200 as vb
    lm = vb.LevenbergMarquardt.new().with_stepbound(1.0).with_patience(1000).with_tol(1e-10).with_scale_diag(False)
    o = lm._o
    assert (o.stepbound, o.patience, o.ftol, o.xtol, o.gtol, o.scale_diag) == (1.0, 1000, 1e-10, 1e-10, 1e-10, 0)
    t = vb.TerminationReason(5)
    assert t.was_successful() and repr(t) == "Converged{xtol}" and not vb.TerminationReason(8).was_successful()


def test_closure_models_validate_like_the_reference_builder():
    """Host closures with partial derivatives (src/model/builder/mod.rs:338-440): a missing derivative, a
    derivative for a parameter the function does not take, or a duplicate derivative are build errors;
    a complete closure model evaluates Phi and the non-zero derivative columns in (function, slot) order."""
    import varpro_b200 as vb
    x = np.linspace(0.0, 2.0, 5)

    def f(x, tau):
        return np.exp(-x / tau)

    def df(x, tau):
        return np.exp(-x / tau) * x / tau ** 2

    with pytest.raises(vb.ModelBuildError):      # missing derivative
        vb.SeparableModelBuilder(["tau"]).function(["tau"], f).independent_variable(x).initial_parameters([1.0]).build()
    with pytest.raises(vb.ModelBuildError):      # derivative for a foreign parameter
        (vb.SeparableModelBuilder(["tau", "w"]).function(["tau"], f).partial_deriv("w", df)
         .function(["w"], vb.ExpDecay()).independent_variable(x).initial_parameters([1.0, 2.0]).build())
    with pytest.raises(vb.ModelBuildError):      # duplicate derivative
        (vb.SeparableModelBuilder(["tau"]).function(["tau"], f).partial_deriv("tau", df).partial_deriv("tau", df)
         .independent_variable(x).initial_parameters([1.0]).build())
    model = (vb.SeparableModelBuilder(["a", "b"])
             .function(["b"], f).partial_deriv("b", df)            # function parameter order need not follow the model's
             .function(["a", "b"], vb.ExpRateCos())
             .invariant_function(lambda x: 2.0 * x)
             .independent_variable(x).initial_parameters([0.5, 1.5]).build())
    assert model.is_host_evaluated()
    assert model.derivative_index() == [(0, 1), (1, 0), (1, 1)]
    cols, dcols = model.eval_host(np.array([0.5, 1.5]))
    assert len(cols) == 3 and len(dcols) == 3
    assert np.allclose(cols[0], np.exp(-x / 1.5)) and np.allclose(cols[2], 2.0 * x)
    assert np.allclose(cols[1], np.exp(-0.5 * x) * np.cos(1.5 * x))
    assert np.allclose(dcols[1], -x * np.exp(-0.5 * x) * np.cos(1.5 * x))      # d/da
    assert np.allclose(dcols[2], -x * np.exp(-0.5 * x) * np.sin(1.5 * x))      # d/db


def test_fit_statistics_accessors_follow_the_reference_ordering():
    """FitStatistics (src/statistics/mod.rs): covariance ordered (c..., alpha...), variances, correlation."""
    import varpro_b200 as vb
    cov = np.array([[4.0, 1.0, 0.5], [1.0, 9.0, -0.3], [0.5, -0.3, 1.0]])
    st = vb.FitStatistics(cov, 0.25, 17, 2, np.array([1.0, 2.0]))
    assert np.allclose(st.linear_coefficients_variance(), [4.0, 9.0])
    assert np.allclose(st.nonlinear_parameters_variance(), [1.0])
    assert abs(st.regression_standard_error() - 0.5) < 1e-15
    corr = st.calculate_correlation_matrix()
    assert np.allclose(np.diag(corr), 1.0) and abs(corr[0, 1] - 1.0 / 6.0) < 1e-15
    from scipy import stats
    assert np.allclose(st.confidence_band_radius(0.9), np.array([1.0, 2.0]) * stats.t.ppf(0.95, 17))
    with pytest.raises(vb.VarproError):
        st.confidence_band_radius(1.5)
