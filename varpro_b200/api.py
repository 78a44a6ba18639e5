"""Host-side mirror of the reference's operator interface for the VarPro hot path.

Names, argument meaning and error behaviour follow geo-ant/varpro v0.13.3 so that
the parity tests read like the reference's own tests:

  SeparableModelBuilder   src/model/builder/mod.rs:252-525
  SeparableModel          src/model/mod.rs:367-517   (trait SeparableNonlinearModel :239-363)
  SeparableProblemBuilder src/problem/builder.rs:116-324
  SeparableProblem        src/problem.rs:57-213 + impl LeastSquaresProblem src/solvers/levmar/mod.rs:22-202
  LevenbergMarquardt      levenberg-marquardt 0.14 knobs (call sites src/solvers/levmar/mod.rs:221,307-315)
  LevMarSolver, FitResult src/solvers/levmar/mod.rs:208-315, src/fit.rs:15-123

Everything numeric happens behind the C ABI (include/varpro_b200.h) in the CUDA
library; this module only validates, marshals and owns handles. The one
deliberate difference from the reference: basis functions are not closures but
descriptors of built-in device functions (a kernel cannot call a host closure;
see DESIGN.md), so `.function(params, ExpDecay())` replaces
`.function(params, exp_decay).partial_deriv(p, exp_decay_dtau)`.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import List, Optional, Sequence

import numpy as np

from . import _lib
from ._lib import VP_F32, VP_F64

# ---------------------------------------------------------------------------
# errors (mirroring the reference's enums)
# ---------------------------------------------------------------------------


class VarproError(Exception):
    """Base class. `.status` holds the vp_status code when it came through the ABI."""

    def __init__(self, msg, status=None):
        super().__init__(msg)
        self.status = status


class ModelBuildError(VarproError):  # src/model/builder/error.rs:5-129
    pass


class DuplicateParameterNames(ModelBuildError):
    pass


class EmptyParameters(ModelBuildError):
    pass


class FunctionParameterNotInModel(ModelBuildError):
    pass


class EmptyModel(ModelBuildError):
    pass


class UnusedParameter(ModelBuildError):
    pass


class IncorrectParameterCount(ModelBuildError):
    pass


class CommaInParameterNameNotAllowed(ModelBuildError):
    pass


class MissingX(ModelBuildError):
    pass


class MissingInitialParameters(ModelBuildError):
    pass


class SeparableProblemBuilderError(VarproError):  # src/problem/builder.rs:15-46
    pass


class YDataMissing(SeparableProblemBuilderError):
    pass


class InvalidLengthOfData(SeparableProblemBuilderError):
    pass


class ZeroLengthVector(SeparableProblemBuilderError):
    pass


class InvalidParameterCount(SeparableProblemBuilderError):
    pass


class InvalidLengthOfWeights(SeparableProblemBuilderError):
    pass


class ModelError(VarproError):  # src/model/errors.rs:5-42
    pass


_STATUS_TO_EXC = {
    1: YDataMissing, 2: InvalidLengthOfData, 3: ZeroLengthVector, 4: InvalidParameterCount,
    5: InvalidLengthOfWeights, 10: FunctionParameterNotInModel, 12: IncorrectParameterCount,
    13: EmptyModel, 14: UnusedParameter,
}


def _check(status: int, ctx=None):
    if status == 0:
        return
    lib = _lib.load()
    msg = lib.vp_last_error(ctx).decode() if ctx is not None else ""
    if not msg:
        msg = lib.vp_status_string(status).decode()
    raise _STATUS_TO_EXC.get(status, VarproError)(msg, status)


# ---------------------------------------------------------------------------
# basis-function descriptors
# ---------------------------------------------------------------------------
@dataclass(frozen=True)
class BasisFunction:
    kind: int
    arity: int
    scale: float = 1.0


def ExpDecay():
    """exp(-x/tau), d/dtau = exp(-x/tau) x/tau^2 (shared_test_code/src/lib.rs:101-114)."""
    return BasisFunction(0, 1)


def Constant():
    """1 -- an invariant function (shared_test_code/src/lib.rs:123)."""
    return BasisFunction(1, 0)


def ExpRateCos():
    """exp(-a x) cos(b x), parameters (a, b) (shared_test_code/src/models.rs:321-322)."""
    return BasisFunction(2, 2)


def SinPhase():
    """sin(omega x + phi), parameters (omega, phi) (src/test_helpers/mod.rs:27-51)."""
    return BasisFunction(3, 2)


def LinearX(scale=1.0):
    """scale * x -- invariant (src/model/builder/test.rs:97,101)."""
    return BasisFunction(4, 0, float(scale))


class HostFunction:
    """A basis function given as a host closure f(x, *params) -> array of len(x), with its partial
    derivatives (the reference's `.function(params, f).partial_deriv(name, df)`,
    src/model/builder/mod.rs:338-440). A model that contains one is evaluated on the host once per
    evaluation (vp_model_create_hosteval); the O(m*S) work stays on the GPU."""

    def __init__(self, f, param_names):
        self.f = f
        self.param_names = list(param_names)
        self.arity = len(self.param_names)
        self.derivs = {}
        self.kind, self.scale = 100, 1.0


def _builtin_numpy(fn: "BasisFunction", x, p):
    """numpy restatement of the built-in device kinds (used when a model mixes them with host closures)."""
    k = fn.kind
    if k == 0:
        e = np.exp(-x / p[0])
        return e, [e * x / (p[0] * p[0])]
    if k == 1:
        return np.ones_like(x), []
    if k == 2:
        e, c, sn = np.exp(-p[0] * x), np.cos(p[1] * x), np.sin(p[1] * x)
        return e * c, [-x * (e * c), -x * e * sn]
    if k == 3:
        return np.sin(p[0] * x + p[1]), [x * np.cos(p[0] * x + p[1]), np.cos(p[0] * x + p[1])]
    if k == 4:
        return fn.scale * x, []
    raise ModelError(f"unsupported basis kind {k}")


# ---------------------------------------------------------------------------
# device context (one per device ordinal, created lazily)
# ---------------------------------------------------------------------------
class _Ctx:
    _by_device = {}
    _lock = __import__("threading").Lock()

    def __init__(self, device: int):
        lib = _lib.load()
        h = C.c_void_p()
        st = lib.vp_ctx_create(device, C.byref(h))
        if st != 0:
            raise VarproError(lib.vp_last_error(None).decode() or lib.vp_status_string(st).decode(), st)
        self.h = h
        self.device = device

    @classmethod
    def get(cls, device: int = 0, slot: int = 0) -> "_Ctx":
        """One library context (= one CUDA stream, used by one host thread at a time) per
        (device, slot). Host threads that want their copies and fits to overlap use different slots."""
        with cls._lock:
            if (device, slot) not in cls._by_device:
                cls._by_device[(device, slot)] = _Ctx(device)
            return cls._by_device[(device, slot)]

    def kernel_launches(self) -> int:
        return int(_lib.load().vp_ctx_kernel_launches(self.h))

    def set_option(self, key: str, value) -> None:
        """vp_ctx_set_option: tunables of this context (see include/varpro_b200.h)."""
        _check(_lib.load().vp_ctx_set_option(self.h, str(key).encode(), str(value).encode()), self.h)

    def trim(self) -> None:
        """vp_ctx_trim: hand the idle cached buffers back to the CUDA allocator."""
        _check(_lib.load().vp_ctx_trim(self.h), self.h)

    def stream(self) -> int:
        return int(_lib.load().vp_ctx_stream(self.h) or 0)


def set_option(key: str, value, device: int = 0, slot: int = 0) -> None:
    """Set a tunable of the (device, slot) context, e.g. set_option("fit_mode", "host")."""
    _Ctx.get(device, slot).set_option(key, value)


def kernel_launches(device: int = 0) -> int:
    """Kernels launched so far by the library on `device` (bench.py's gpu_launches)."""
    return _Ctx.get(device).kernel_launches()


# ---------------------------------------------------------------------------
# model
# ---------------------------------------------------------------------------
class SeparableModel:
    """Result of SeparableModelBuilder.build (src/model/mod.rs:367-517)."""

    def __init__(self, parameter_names, functions, x, initial_parameters, dtype):
        self._names = list(parameter_names)
        self._functions = functions  # list of (BasisFunction, [param indices])
        self.x = np.ascontiguousarray(x, dtype=dtype)
        self._params = np.array(initial_parameters, dtype=np.float64)
        self.dtype = np.dtype(dtype)

    def parameters(self) -> List[str]:
        return list(self._names)

    def parameter_count(self) -> int:
        return len(self._names)

    def base_function_count(self) -> int:
        return len(self._functions)

    def output_len(self) -> int:
        return int(self.x.shape[0])

    def params(self) -> np.ndarray:
        return self._params.copy()

    def set_params(self, parameters):
        p = np.asarray(parameters, dtype=np.float64).ravel()
        if p.shape[0] != self.parameter_count():
            raise ModelError(f"Model expects {self.parameter_count()} parameters, but got {p.shape[0]}")
        self._params = p.copy()

    def is_host_evaluated(self) -> bool:
        return any(isinstance(f, HostFunction) for f, _ in self._functions)

    def derivative_index(self):
        """(basis j, parameter k) of every non-zero derivative column, ordered by (function, slot)."""
        return [(j, k) for j, (_, idx) in enumerate(self._functions) for k in idx]

    def eval_host(self, alpha):
        """model.eval() and the non-zero columns of eval_partial_deriv(k) on the host (f64)."""
        x = np.asarray(self.x, dtype=np.float64)
        cols, dcols = [], []
        for f, idx in self._functions:
            p = [alpha[i] for i in idx]
            if isinstance(f, HostFunction):
                v = np.asarray(f.f(x, *p), dtype=np.float64)
                ds = [np.asarray(f.derivs[nm](x, *p), dtype=np.float64) for nm in f.param_names]
            else:
                v, ds = _builtin_numpy(f, x, p)
            if v.shape != x.shape or any(d.shape != x.shape for d in ds):
                raise ModelError(f"Basis function gave vector of length {v.shape}, but expected output length {x.shape}")
            cols.append(v)
            dcols.extend(ds)
        return cols, dcols

    def _descs(self):
        arr = (_lib.BasisDesc * len(self._functions))()
        for d, (f, idx) in zip(arr, self._functions):
            d.kind, d.n_params, d.scale = f.kind, len(idx), f.scale
            for i, v in enumerate(idx):
                d.param_idx[i] = v
        return arr


class SeparableModelBuilder:
    """State machine of src/model/builder/mod.rs:252-525. Errors are deferred to build()."""

    def __init__(self, parameter_names: Sequence[str], dtype=np.float64):
        self._names = list(parameter_names)
        self._functions = []
        self._x = None
        self._p0 = None
        self._dtype = np.dtype(dtype)
        self._error: Optional[ModelBuildError] = None
        self._check_names(self._names)

    @classmethod
    def new(cls, parameter_names, dtype=np.float64):
        return cls(parameter_names, dtype)

    def _fail(self, err):
        if self._error is None:
            self._error = err

    def _check_names(self, names):
        # check_parameter_names (src/model/detail.rs:14-46)
        if len(names) == 0:
            return self._fail(EmptyParameters(
                "A function or model parameter list is empty! It must at least contain one parameter."))
        for nm in names:
            if "," in nm:
                return self._fail(CommaInParameterNameNotAllowed(
                    f"Parameter names may not contain comma separator: '{nm}'."))
        if len(set(names)) != len(names):
            return self._fail(DuplicateParameterNames(f"Parameter list {list(names)} contains duplicates!"))

    def function(self, function_params: Sequence[str], function):
        """`function`: a built-in device basis function (ExpDecay(), ...) or any host callable
        f(x, *params) -> array, whose derivatives are then added with .partial_deriv(name, df)."""
        fp = list(function_params)
        self._check_names(fp)
        if self._error:
            return self
        if not isinstance(function, BasisFunction):
            if not callable(function):
                self._fail(ModelBuildError("function must be a BasisFunction or a callable"))
                return self
            function = HostFunction(function, fp)
        if len(fp) != function.arity:
            self._fail(IncorrectParameterCount(
                f"Incorrect number of parameters for function: expected {function.arity}, got {len(fp)}"))
            return self
        idx = []
        for nm in fp:  # create_index_mapping (src/model/detail.rs:60-78)
            if nm not in self._names:
                self._fail(FunctionParameterNotInModel(
                    f"Function parameter '{nm}' is not part of the model parameters."))
                return self
            idx.append(self._names.index(nm))
        self._functions.append((function, idx))
        return self

    def invariant_function(self, function):
        if not isinstance(function, BasisFunction) and callable(function):
            function = HostFunction(function, [])
        if function.arity != 0:
            self._fail(IncorrectParameterCount(
                f"Incorrect number of parameters for function: expected {function.arity}, got 0"))
            return self
        self._functions.append((function, []))
        return self

    def partial_deriv(self, parameter: str, derivative=None):
        """src/model/builder/mod.rs:387-440. For the built-in device kinds the derivatives are built in
        (accepted and ignored); for a host closure the derivative callable df(x, *params) is required."""
        if self._error or not self._functions:
            return self
        f, _ = self._functions[-1]
        if isinstance(f, HostFunction):
            if parameter not in f.param_names:
                self._fail(FunctionParameterNotInModel(
                    f"Function parameter '{parameter}' is not part of the function's parameters."))
            elif parameter in f.derivs:
                self._fail(ModelBuildError(f"Derivative for parameter '{parameter}' was already provided!"))
            elif not callable(derivative):
                self._fail(ModelBuildError("partial_deriv of a host function needs a callable"))
            else:
                f.derivs[parameter] = derivative
        return self

    def independent_variable(self, x):
        self._x = np.asarray(x)
        return self

    def initial_parameters(self, initial_parameters):
        self._p0 = list(initial_parameters)
        return self

    def build(self) -> SeparableModel:
        if self._error:
            raise self._error
        if not self._functions:
            raise EmptyModel("Tried to construct model with no functions. A model must contain at least one function.")
        used = {i for _, idx in self._functions for i in idx}
        for k, nm in enumerate(self._names):  # src/model/builder/mod.rs:547-557
            if k not in used:
                raise UnusedParameter(f"Model depends on parameter '{nm}', but none of its functions use it.")
        for f, _ in self._functions:  # MissingDerivative (src/model/builder/error.rs)
            if isinstance(f, HostFunction):
                for nm in f.param_names:
                    if nm not in f.derivs:
                        raise ModelBuildError(f"Missing partial derivative for parameter '{nm}'.")
        if self._x is None:
            raise MissingX("Missing vector for independent variable x")
        if self._p0 is None:
            raise MissingInitialParameters("Missing initial guesses for model parameters")
        if len(self._p0) != len(self._names):
            raise IncorrectParameterCount(
                f"Incorrect number of parameters for function: expected {len(self._names)}, got {len(self._p0)}")
        return SeparableModel(self._names, self._functions, self._x, self._p0, self._dtype)


# ---------------------------------------------------------------------------
# problem
# ---------------------------------------------------------------------------
class SeparableProblem:
    """Device-resident SeparableProblem (src/problem.rs:57-107). Created by the builder."""

    def __init__(self, model: SeparableModel, Y: np.ndarray, weights, eps, single_rhs, device=0,
                 y_device_ptr=None, S=None, ldY=None, ctx_slot=0):
        lib = _lib.load()
        self._ctx = _Ctx.get(device, ctx_slot)
        self.model_host = model
        self.single_rhs = single_rhs
        self.dtype = model.dtype
        self._m = model.output_len()
        self._n = model.base_function_count()
        self._q = model.parameter_count()
        self._model_h = C.c_void_p()
        self._h = C.c_void_p()
        dt = VP_F32 if self.dtype == np.float32 else VP_F64
        if model.is_host_evaluated():
            ind = model.derivative_index()
            p = len(ind)
            ind_arr = (C.c_int32 * max(2 * p, 1))(*[v for jk in ind for v in jk])
            m_, n_ = self._m, self._n

            def _cb(_user, alpha_p, phi_p, dphi_p):
                try:
                    alpha = np.array([alpha_p[i] for i in range(self._q)], dtype=np.float64)
                    cols, dcols = model.eval_host(alpha)
                    phi = np.ctypeslib.as_array(phi_p, shape=(n_, m_))      # column-major m x n == row-major n x m
                    for j, c in enumerate(cols):
                        phi[j, :] = c
                    if p:
                        dphi = np.ctypeslib.as_array(dphi_p, shape=(p, m_))
                        for e, c in enumerate(dcols):
                            dphi[e, :] = c
                    return 0
                except Exception:  # a model error: the cache becomes None (src/solvers/levmar/mod.rs:43-45)
                    return 1

            self._host_cb = _lib.HOST_EVAL_FN(_cb)  # must outlive the model handle
            _check(lib.vp_model_create_hosteval(self._ctx.h, dt, self._m, self._q, self._n, p, ind_arr,
                                                C.cast(self._host_cb, C.c_void_p), None, C.byref(self._model_h)),
                   self._ctx.h)
        else:
            descs = model._descs()
            _check(lib.vp_model_create(self._ctx.h, dt, self._m, model.x.ctypes.data_as(C.c_void_p), self._q,
                                       self._n, descs, C.byref(self._model_h)), self._ctx.h)
        a0 = np.ascontiguousarray(model.params(), dtype=np.float64)
        w = None if weights is None else np.ascontiguousarray(weights, dtype=self.dtype)
        wp = None if w is None else w.ctypes.data_as(C.c_void_p)
        ap = a0.ctypes.data_as(C.POINTER(C.c_double))
        try:
            if y_device_ptr is not None:
                self._S = int(S)
                _check(lib.vp_problem_create_device(self._ctx.h, self._model_h, self._S, C.c_void_p(y_device_ptr),
                                                    int(ldY), wp, float(eps), ap, C.byref(self._h)), self._ctx.h)
            else:
                self._S = int(Y.shape[1])
                _check(lib.vp_problem_create(self._ctx.h, self._model_h, self._S, Y.ctypes.data_as(C.c_void_p),
                                             int(Y.shape[0]), wp, float(eps), ap, C.byref(self._h)), self._ctx.h)
        except Exception:
            lib.vp_model_destroy(self._model_h)
            self._model_h = C.c_void_p()
            raise
        self._weights = w

    def close(self):
        lib = _lib.load()
        if getattr(self, "_h", None) and self._h.value:
            lib.vp_problem_destroy(self._h)
            self._h = C.c_void_p()
        if getattr(self, "_model_h", None) and self._model_h.value:
            lib.vp_model_destroy(self._model_h)
            self._model_h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # --- impl LeastSquaresProblem (src/solvers/levmar/mod.rs:22-202) ---
    def set_params(self, params):
        a = np.ascontiguousarray(params, dtype=np.float64)
        if a.shape[0] != self._q:
            raise ModelError(f"Model expects {self._q} parameters, but got {a.shape[0]}")
        _check(_lib.load().vp_set_params(self._h, a.ctypes.data_as(C.POINTER(C.c_double))), self._ctx.h)
        self.model_host._params = a.copy()

    def params(self) -> np.ndarray:
        out = np.empty(self._q, dtype=np.float64)
        _check(_lib.load().vp_params(self._h, out.ctypes.data_as(C.POINTER(C.c_double))), self._ctx.h)
        return out

    def _fetch(self, fn, shape):
        out = np.empty(shape, dtype=self.dtype, order="F")
        st = fn(self._h, out.ctypes.data_as(C.c_void_p))
        if st == 20:  # cache is None -> Option::None in the reference
            return None
        _check(st, self._ctx.h)
        return out

    def residuals(self) -> Optional[np.ndarray]:
        return self._fetch(_lib.load().vp_residuals, (self._m * self._S,))

    def jacobian(self) -> Optional[np.ndarray]:
        return self._fetch(_lib.load().vp_jacobian, (self._m * self._S, self._q))

    def linear_coefficients(self) -> Optional[np.ndarray]:
        c = self._fetch(_lib.load().vp_linear_coefficients, (self._n, self._S))
        if c is not None and self.single_rhs:
            return c[:, 0]
        return c

    def best_fit(self) -> Optional[np.ndarray]:
        b = self._fetch(_lib.load().vp_best_fit, (self._m, self._S))
        if b is not None and self.single_rhs:
            return b[:, 0]
        return b

    def reduce(self):
        """(||r||^2, J^T r, J^T J) of the current parameters, or None if the cache is None."""
        r = _lib.Reduced()
        st = _lib.load().vp_reduce(self._h, C.byref(r))
        if st == 20:
            return None
        _check(st, self._ctx.h)
        q = self._q
        return dict(rnorm2=r.rnorm2, g=np.array(r.g[:q]), H=np.array(r.H[:q * q]).reshape(q, q).T.copy(),
                    finite=bool(r.finite))

    def set_jacobian(self, mode: str):
        """"kaufman" (default; what the reference implements) or "full" (adds the second Golub-Pereyra term
        the reference leaves as a TODO, src/solvers/levmar/mod.rs:188-190)."""
        code = {"kaufman": 0, "full": 1}[mode]
        _check(_lib.load().vp_problem_set_jacobian(self._h, code), self._ctx.h)
        return self

    def set_rank_policy(self, policy: str):
        """"absolute" (default; the reference: singular values <= epsilon are truncated in the solve,
        src/solvers/levmar/mod.rs:52-54) or "relative" (MATLAB: sigma <= m*eps*sigma_1, matlab/varpro.m:642-643)."""
        _check(_lib.load().vp_problem_set_rank_policy(self._h, {"absolute": 0, "relative": 1}[policy]), self._ctx.h)
        return self

    def statistics(self, confidence_sigma: bool = False):
        """FitStatistics::try_calculate (src/statistics/mod.rs:352-441) for every right-hand side with the
        shared nonlinear parameters (vp_statistics). Returns a list of FitStatistics (one per column)."""
        t = self._n + self._q
        cov = np.empty((self._S, t, t), dtype=np.float64)
        chi2 = np.empty(self._S, dtype=np.float64)
        conf = np.empty((self._S, self._m), dtype=np.float64) if confidence_sigma else None
        dp = C.POINTER(C.c_double)
        _check(_lib.load().vp_statistics(self._h, cov.ctypes.data_as(dp), chi2.ctypes.data_as(dp),
                                         conf.ctypes.data_as(dp) if conf is not None else None), self._ctx.h)
        dof = self._m - t
        return [FitStatistics(cov[s].T.copy(), float(chi2[s]), dof, self._n, None if conf is None else conf[s])
                for s in range(self._S)]

    def model(self) -> SeparableModel:
        self.model_host._params = self.params()  # the fitted parameters live in the library
        return self.model_host

    def weights(self):
        return self._weights


class SeparableProblemBuilder:
    """src/problem/builder.rs:116-324."""

    def __init__(self, model: SeparableModel, single_rhs: bool):
        self._model = model
        self._single = single_rhs
        self._Y = None
        self._w = None
        self._eps = None
        self._device = 0

    @classmethod
    def new(cls, model):  # :116
        return cls(model, True)

    @classmethod
    def mrhs(cls, model):  # :194
        return cls(model, False)

    def observations(self, observed):  # :142 / :220
        y = np.asarray(observed)
        if self._single:
            y = y.reshape(-1, 1) if y.ndim == 1 else y
        self._Y = y
        return self

    def weights(self, weights):  # :261-266
        self._w = np.asarray(weights)
        return self

    def epsilon(self, eps):  # :246-251
        self._eps = abs(float(eps))
        return self

    def device(self, ordinal: int, ctx_slot: int = 0):
        self._device = int(ordinal)
        self._ctx_slot = int(ctx_slot)
        return self

    def build(self) -> SeparableProblem:  # :278-324
        if self._Y is None:
            raise YDataMissing("Right hand side(s) not provided", 1)
        Y = self._Y
        x_len = self._model.output_len()
        if x_len == 0 or Y.size == 0:
            raise ZeroLengthVector("x or y must have nonzero number of elements.", 3)
        if Y.ndim != 2 or (self._single and Y.shape[1] != 1):
            raise InvalidLengthOfData("observations have the wrong shape", 2)
        if x_len != Y.shape[0]:
            raise InvalidLengthOfData(
                f"Vectors x and y must have same lengths. Given x length = {x_len} and y length = {Y.shape[0]}", 2)
        if self._w is not None and self._w.shape != (Y.shape[0],):
            raise InvalidLengthOfWeights("The weights must have the same length as the data y.", 5)
        eps = -1.0 if self._eps is None else self._eps  # default: machine epsilon of the scalar (:282)
        Yf = np.asfortranarray(Y, dtype=self._model.dtype)
        return SeparableProblem(self._model, Yf, self._w, eps, self._single, self._device,
                                ctx_slot=getattr(self, "_ctx_slot", 0))


# ---------------------------------------------------------------------------
# solver
# ---------------------------------------------------------------------------
TERMINATION_NAMES = [
    "User", "Numerical", "ResidualsZero", "Orthogonal", "Converged{ftol}", "Converged{xtol}",
    "Converged{ftol,xtol}", "NoImprovementPossible", "LostPatience", "NoParameters", "NoResiduals",
    "WrongDimensions",
]


class TerminationReason:
    def __init__(self, code: int):
        self.code = code

    def was_successful(self) -> bool:
        return self.code in (2, 3, 4, 5, 6)

    def __repr__(self):
        return TERMINATION_NAMES[self.code] if 0 <= self.code < len(TERMINATION_NAMES) else f"?({self.code})"


@dataclass
class MinimizationReport:
    termination: TerminationReason
    number_of_evaluations: int
    objective_function: float


class LevenbergMarquardt:
    """Option holder with the levenberg-marquardt crate's builder knobs. Unset knobs are passed as -1 (= crate
    default: ftol = xtol = gtol = 30 eps, stepbound 100, patience 100, scale_diag on); a tolerance of 0 is a legal
    value and disables that criterion, as in the crate."""

    def __init__(self):
        self._o = _lib.LmOptions(-1.0, -1.0, -1.0, -1.0, -1, -1)

    @classmethod
    def new(cls):
        return cls()

    def with_ftol(self, v):
        self._o.ftol = float(v)
        return self

    def with_xtol(self, v):
        self._o.xtol = float(v)
        return self

    def with_gtol(self, v):
        self._o.gtol = float(v)
        return self

    def with_tol(self, v):
        return self.with_ftol(v).with_xtol(v).with_gtol(v)

    def with_stepbound(self, v):
        self._o.stepbound = float(v)
        return self

    def with_patience(self, v):
        self._o.patience = int(v)
        return self

    def with_scale_diag(self, v):
        self._o.scale_diag = 1 if v else 0
        return self


class FitResult:
    """src/fit.rs:15-123."""

    def __init__(self, problem: SeparableProblem, report: MinimizationReport):
        self.problem = problem
        self.minimization_report = report

    def nonlinear_parameters(self) -> np.ndarray:
        return self.problem.params()

    def linear_coefficients(self):
        return self.problem.linear_coefficients()

    def best_fit(self):
        return self.problem.best_fit()

    def was_successful(self) -> bool:
        return self.minimization_report.termination.was_successful()


class FitStatistics:
    """src/statistics/mod.rs:24-345: covariance ordered (c..., alpha...), reduced chi^2, confidence band."""

    def __init__(self, covariance, reduced_chi2, degrees_of_freedom, n_linear, conf_sigma):
        self._cov = covariance
        self._chi2 = reduced_chi2
        self._dof = degrees_of_freedom
        self._n = n_linear
        self._sigma = conf_sigma

    def covariance_matrix(self) -> np.ndarray:  # :66-76
        return self._cov

    def reduced_chi2(self) -> float:  # :244-246
        return self._chi2

    def regression_standard_error(self) -> float:  # :250-252
        return float(np.sqrt(self._chi2))

    def linear_coefficients_variance(self) -> np.ndarray:  # :134-139
        return np.diag(self._cov)[: self._n].copy()

    def nonlinear_parameters_variance(self) -> np.ndarray:  # :127-132
        return np.diag(self._cov)[self._n:].copy()

    def calculate_correlation_matrix(self) -> np.ndarray:  # :446-472
        d = np.sqrt(np.diag(self._cov))
        return self._cov / np.outer(d, d)

    def confidence_band_radius(self, probability: float) -> np.ndarray:  # :275-291
        if not (0.0 < probability < 1.0):
            raise VarproError("probability must be in (0, 1)")
        if self._sigma is None:
            raise VarproError("statistics were computed without confidence_sigma=True")
        from scipy import stats
        return self._sigma * stats.t.ppf((probability + 1.0) / 2.0, self._dof)


class FitError(VarproError):
    """Err(FitResult) of LevMarSolver::fit: carries the same FitResult in `.result`."""

    def __init__(self, result: FitResult):
        super().__init__(f"fit did not terminate successfully: {result.minimization_report.termination!r}")
        self.result = result


class LevMarSolver:
    """src/solvers/levmar/mod.rs:208-315."""

    def __init__(self, solver: Optional[LevenbergMarquardt] = None):
        self._solver = solver or LevenbergMarquardt()

    @classmethod
    def default(cls):
        return cls()

    @classmethod
    def with_solver(cls, solver: LevenbergMarquardt):
        return cls(solver)

    def fit(self, problem: SeparableProblem) -> FitResult:
        """Ok(FitResult) is returned, Err(FitResult) is raised as FitError (same payload)."""
        rep = _lib.FitReport()
        _check(_lib.load().vp_fit(problem._h, C.byref(self._solver._o), C.byref(rep)), problem._ctx.h)
        problem.model_host._params = problem.params()
        result = FitResult(problem, MinimizationReport(TerminationReason(rep.termination),
                                                       rep.number_of_evaluations, rep.objective_function))
        if not result.was_successful():
            raise FitError(result)
        return result

    def fit_with_statistics(self, problem: SeparableProblem):
        """src/solvers/levmar/mod.rs:275-304. The reference allows SingleRhs only; for an MRHS problem the
        statistics of every column are returned as a list."""
        result = self.fit(problem)
        st = problem.statistics(confidence_sigma=True)
        return result, (st[0] if problem.single_rhs else st)

    def fit_host_batch(self, model: "SeparableModel", observations: Sequence[np.ndarray], weights=None, eps: float = -1.0,
                       device: int = 0, workers: int = 0):
        """Build, fit and read back one MRHS problem per entry of `observations` (m x S arrays, Fortran order; pinned
        host memory gives full PCIe speed) through vp_fit_host_batch: worker threads of the library pipeline the
        host-to-device copy of one problem with the fit of another. Equivalent to
        `[LevMarSolver.fit(SeparableProblemBuilder.mrhs(model).observations(Y).build()) for Y in observations]`.
        Returns (reports, parameters (len x q), coefficients list of n x S arrays)."""
        lib = _lib.load()
        ctx = _Ctx.get(device)
        Ys = [Y if (isinstance(Y, np.ndarray) and Y.flags.f_contiguous and Y.dtype == model.dtype) else np.asfortranarray(Y, dtype=model.dtype)
              for Y in observations]
        if not Ys:
            return [], np.empty((0, model.parameter_count())), []
        m, S = Ys[0].shape
        if any(Y.shape != (m, S) for Y in Ys) or m != model.output_len():
            raise InvalidLengthOfData(f"Vectors x and y must have same lengths. Given x length = {model.output_len()} and y length = {m}")
        npb, q, n = len(Ys), model.parameter_count(), model.base_function_count()
        Cs = [np.empty((n, S), dtype=model.dtype, order="F") for _ in range(npb)]
        alpha = np.empty((npb, q), dtype=np.float64)
        reps = (_lib.FitReport * npb)()
        yptrs = (C.c_void_p * npb)(*[Y.ctypes.data for Y in Ys])
        cptrs = (C.c_void_p * npb)(*[c.ctypes.data for c in Cs])
        w = None if weights is None else np.ascontiguousarray(weights, dtype=model.dtype)
        a0 = np.ascontiguousarray(model.params(), dtype=np.float64)
        _check(lib.vp_fit_host_batch(ctx.h, VP_F32 if model.dtype == np.float32 else VP_F64, m, model.x.ctypes.data_as(C.c_void_p), q, n,
                                     model._descs(), npb, S, yptrs, m, None if w is None else w.ctypes.data_as(C.c_void_p),
                                     float(eps), a0.ctypes.data_as(C.POINTER(C.c_double)), C.byref(self._solver._o), int(workers),
                                     reps, alpha.ctypes.data_as(C.POINTER(C.c_double)), cptrs), ctx.h)
        reports = [MinimizationReport(TerminationReason(r.termination), r.number_of_evaluations, r.objective_function) for r in reps]
        return reports, alpha, Cs

    def fit_many(self, problems: Sequence[SeparableProblem], max_concurrent: int = 0) -> List[FitResult]:
        """Fit independent problems together (vp_fit_many): the loop a caller of the reference writes around
        LevMarSolver::fit, executed by ONE persistent kernel whose CTAs take (fit, group-of-columns) work items
        from a device-side queue. Results are bitwise those of `fit` per problem. Returns one FitResult per
        problem in order; unsuccessful fits are returned (not raised) -- check `was_successful()` as with the
        reference's Err(FitResult). `max_concurrent` is ignored (kept for callers of the first ABI version)."""
        problems = list(problems)
        if not problems:
            return []
        n = len(problems)
        handles = (C.c_void_p * n)(*[p._h for p in problems])
        reps = (_lib.FitReport * n)()
        _check(_lib.load().vp_fit_many(handles, n, C.byref(self._solver._o), reps, int(max_concurrent)),
               problems[0]._ctx.h)
        # (no per-problem calls here: the host-side model parameters are refreshed lazily by model())
        return [FitResult(p, MinimizationReport(TerminationReason(rep.termination), rep.number_of_evaluations,
                                                rep.objective_function)) for p, rep in zip(problems, reps)]


# ---------------------------------------------------------------------------
# independent batch (BASELINE config 3): P single-RHS problems in one launch
# ---------------------------------------------------------------------------
class BatchFitResult:
    """Per-problem results of IndependentBatch.fit: arrays over the P problems."""

    def __init__(self, params, coefficients, terminations, evaluations, objectives):
        self.nonlinear_parameters = params            # (P, q)
        self.linear_coefficients = coefficients       # (n, P)
        self.terminations = [TerminationReason(int(t)) for t in terminations]
        self.number_of_evaluations = evaluations      # (P,)
        self.objective_function = objectives          # (P,) 0.5*||r_w||^2
        self.successful = np.array([t.was_successful() for t in self.terminations])


class IndependentBatch:
    """P independent problems sharing the model structure, x and weights (vp_batch): the loop
    `for p: LevMarSolver.fit(SeparableProblemBuilder.new(model_p).observations(Y[:, p]).build())`
    over the reference API, as one kernel launch. `initial_parameters`: (P, q)."""

    def __init__(self, model: SeparableModel, Y, initial_parameters, weights=None, eps=-1.0, device=0,
                 y_device_ptr=None, P=None):
        lib = _lib.load()
        if model.dtype != np.float64:
            raise VarproError("IndependentBatch: fp64 models only")
        self._ctx = _Ctx.get(device)
        self._m, self._n, self._q = model.output_len(), model.base_function_count(), model.parameter_count()
        self._model_h = C.c_void_p()
        self._h = C.c_void_p()
        _check(lib.vp_model_create(self._ctx.h, VP_F64, self._m, model.x.ctypes.data_as(C.c_void_p), self._q, self._n,
                                   model._descs(), C.byref(self._model_h)), self._ctx.h)
        a0 = np.ascontiguousarray(initial_parameters, dtype=np.float64)  # (P, q) row-major == q x P column-major
        w = None if weights is None else np.ascontiguousarray(weights, dtype=np.float64)
        wp = None if w is None else w.ctypes.data_as(C.c_void_p)
        try:
            if y_device_ptr is not None:
                self._P = int(P)
                if a0.shape != (self._P, self._q):
                    raise InvalidParameterCount("initial_parameters must have shape (P, q)")
                _check(lib.vp_batch_create_device(self._ctx.h, self._model_h, self._P, C.c_void_p(y_device_ptr), self._m, wp,
                                                  float(eps), a0.ctypes.data_as(C.POINTER(C.c_double)), C.byref(self._h)),
                       self._ctx.h)
            else:
                Yf = np.asfortranarray(Y, dtype=np.float64)
                if Yf.ndim != 2 or Yf.shape[0] != self._m:
                    raise InvalidLengthOfData(f"Vectors x and y must have same lengths. Given x length = {self._m} "
                                              f"and y length = {Yf.shape[0]}")
                self._P = Yf.shape[1]
                if a0.shape != (self._P, self._q):
                    raise InvalidParameterCount("initial_parameters must have shape (P, q)")
                _check(lib.vp_batch_create(self._ctx.h, self._model_h, self._P, Yf.ctypes.data_as(C.c_void_p), self._m, wp,
                                           float(eps), a0.ctypes.data_as(C.POINTER(C.c_double)), C.byref(self._h)),
                       self._ctx.h)
        except Exception:
            lib.vp_model_destroy(self._model_h)
            self._model_h = C.c_void_p()
            raise

    def set_params(self, parameters):
        a = np.ascontiguousarray(parameters, dtype=np.float64)
        if a.shape != (self._P, self._q):
            raise InvalidParameterCount("parameters must have shape (P, q)")
        _check(_lib.load().vp_batch_set_params(self._h, a.ctypes.data_as(C.POINTER(C.c_double))), self._ctx.h)

    def set_rank_policy(self, policy: str):
        _check(_lib.load().vp_batch_set_rank_policy(self._h, {"absolute": 0, "relative": 1}[policy]), self._ctx.h)
        return self

    def fit(self, solver: Optional["LevMarSolver"] = None, reports: bool = True) -> Optional[BatchFitResult]:
        lib = _lib.load()
        opts = (solver or LevMarSolver())._solver._o
        reps = (_lib.FitReport * self._P)() if reports else None
        _check(lib.vp_batch_fit(self._h, C.byref(opts), reps), self._ctx.h)
        if not reports:
            return None
        alpha = np.empty((self._P, self._q), dtype=np.float64)
        Cc = np.empty((self._n, self._P), dtype=np.float64, order="F")
        _check(lib.vp_batch_params(self._h, alpha.ctypes.data_as(C.POINTER(C.c_double))), self._ctx.h)
        _check(lib.vp_batch_linear_coefficients(self._h, Cc.ctypes.data_as(C.POINTER(C.c_double))), self._ctx.h)
        rr = np.frombuffer(reps, dtype=np.dtype([("termination", "<i4"), ("nfev", "<i4"), ("obj", "<f8"),
                                                  ("ok", "<i4"), ("res", "<i4")]))
        return BatchFitResult(alpha, Cc, rr["termination"].copy(), rr["nfev"].copy(), rr["obj"].copy())

    def close(self):
        lib = _lib.load()
        if getattr(self, "_h", None) and self._h.value:
            lib.vp_batch_destroy(self._h)
            self._h = C.c_void_p()
        if getattr(self, "_model_h", None) and self._model_h.value:
            lib.vp_model_destroy(self._model_h)
            self._model_h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
