"""ctypes binding of the CPU oracle (oracle/varpro_oracle.c) -- TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / `--impl reference`
legs may import this module. The product package (varpro_b200) never does.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libvarpro_oracle.so")

EXP_DECAY, CONSTANT, EXP_RATE_COS, SIN_PHASE, LINEAR_X = range(5)

TERMINATION = [
    "User", "Numerical", "ResidualsZero", "Orthogonal", "Converged{ftol}", "Converged{xtol}",
    "Converged{ftol,xtol}", "NoImprovementPossible", "LostPatience", "NoParameters",
    "NoResiduals", "WrongDimensions",
]


class _Basis(C.Structure):
    _fields_ = [("kind", C.c_int), ("n_params", C.c_int), ("param_idx", C.c_int * 4),
                ("scale", C.c_double)]


class _Opts(C.Structure):
    _fields_ = [("ftol", C.c_double), ("xtol", C.c_double), ("gtol", C.c_double),
                ("stepbound", C.c_double), ("patience", C.c_int), ("scale_diag", C.c_int)]


class _Report(C.Structure):
    _fields_ = [("termination", C.c_int), ("number_of_evaluations", C.c_int),
                ("number_of_jacobians", C.c_int), ("objective_function", C.c_double),
                ("successful", C.c_int)]


def build(force: bool = False) -> str:
    """Compile the oracle with the committed Makefile (building the checker is not using it)."""
    if force or not os.path.exists(_LIB_PATH) or (
            os.path.getmtime(_LIB_PATH) < os.path.getmtime(os.path.join(_HERE, "varpro_oracle.c"))):
        subprocess.run(["make", "-C", _HERE], check=True, capture_output=True)
    return _LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB_PATH):
            build()
        L = C.CDLL(_LIB_PATH)
        dp = C.POINTER(C.c_double)
        L.vo_problem_new.restype = C.c_void_p
        L.vo_problem_new.argtypes = [C.c_int, C.c_int, C.c_int, C.POINTER(_Basis), dp, C.c_int, dp,
                                     dp, C.c_double, dp]
        L.vo_problem_free.argtypes = [C.c_void_p]
        L.vo_set_threads.argtypes = [C.c_int]
        for name in ("vo_set_params", "vo_residuals", "vo_jacobian", "vo_linear_coefficients",
                     "vo_model_eval", "vo_best_fit"):
            getattr(L, name).argtypes = [C.c_void_p, dp]
            getattr(L, name).restype = C.c_int
        L.vo_params.argtypes = [C.c_void_p, dp]
        L.vo_model_eval_partial_deriv.argtypes = [C.c_void_p, C.c_int, dp]
        L.vo_fit.argtypes = [C.c_void_p, C.POINTER(_Opts), C.POINTER(_Report)]
        L.vo_statistics.argtypes = [C.c_void_p, C.c_int, dp, dp, dp, dp]
        _lib = L
    return _lib


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def make_basis(specs):
    """specs: list of (kind, [param indices], scale?)"""
    arr = (_Basis * len(specs))()
    for b, spec in zip(arr, specs):
        kind, idx = spec[0], list(spec[1])
        b.kind, b.n_params = kind, len(idx)
        for i, v in enumerate(idx):
            b.param_idx[i] = v
        b.scale = spec[2] if len(spec) > 2 else 1.0
    return arr


class OracleProblem:
    """Mirror of SeparableProblem (src/problem.rs:57-83) backed by the C oracle."""

    def __init__(self, x, basis_specs, q, Y, alpha0, weights=None, eps=np.finfo(np.float64).eps):
        x = np.ascontiguousarray(x, dtype=np.float64)
        Y = np.asarray(Y, dtype=np.float64)
        if Y.ndim == 1:
            Y = Y[:, None]
        self.m, self.S = Y.shape
        self.n, self.q = len(basis_specs), q
        Yf = np.asfortranarray(Y)
        w = None if weights is None else np.ascontiguousarray(weights, dtype=np.float64)
        a0 = np.ascontiguousarray(alpha0, dtype=np.float64)
        self._basis = make_basis(basis_specs)
        self._h = lib().vo_problem_new(self.m, self.n, q, self._basis, _dp(x), self.S, _dp(Yf),
                                       None if w is None else _dp(w), float(eps), _dp(a0))
        if not self._h:
            raise ValueError("oracle: invalid problem dimensions")

    def __del__(self):
        if getattr(self, "_h", None):
            lib().vo_problem_free(self._h)
            self._h = None

    def set_params(self, alpha):
        a = np.ascontiguousarray(alpha, dtype=np.float64)
        return lib().vo_set_params(self._h, _dp(a)) == 0

    def params(self):
        out = np.empty(self.q)
        lib().vo_params(self._h, _dp(out))
        return out

    def residuals(self):
        out = np.empty(self.m * self.S)
        return out if lib().vo_residuals(self._h, _dp(out)) == 0 else None

    def jacobian(self):
        out = np.empty((self.m * self.S, self.q), order="F")
        return out if lib().vo_jacobian(self._h, _dp(out)) == 0 else None

    def linear_coefficients(self):
        out = np.empty((self.n, self.S), order="F")
        return out if lib().vo_linear_coefficients(self._h, _dp(out)) == 0 else None

    def model_eval(self):
        out = np.empty((self.m, self.n), order="F")
        lib().vo_model_eval(self._h, _dp(out))
        return out

    def model_eval_partial_deriv(self, k):
        out = np.empty((self.m, self.n), order="F")
        lib().vo_model_eval_partial_deriv(self._h, k, _dp(out))
        return out

    def best_fit(self):
        out = np.empty((self.m, self.S), order="F")
        return out if lib().vo_best_fit(self._h, _dp(out)) == 0 else None

    def fit(self, ftol=0.0, xtol=0.0, gtol=0.0, stepbound=0.0, patience=0, scale_diag=-1):
        o = _Opts(ftol, xtol, gtol, stepbound, patience, scale_diag)
        r = _Report()
        lib().vo_fit(self._h, C.byref(o), C.byref(r))
        return dict(termination=TERMINATION[r.termination] if r.termination >= 0 else "?",
                    number_of_evaluations=r.number_of_evaluations,
                    number_of_jacobians=r.number_of_jacobians,
                    objective_function=r.objective_function, successful=bool(r.successful))

    def statistics(self, s=0):
        t = self.n + self.q
        cov = np.empty((t, t), order="F")
        chi2 = C.c_double()
        wres = np.empty(self.m)
        conf = np.empty(self.m)
        rc = lib().vo_statistics(self._h, s, _dp(cov), C.byref(chi2), _dp(wres), _dp(conf))
        if rc != 0:
            return None
        return dict(covariance=np.array(cov), reduced_chi2=chi2.value, weighted_residuals=wres,
                    unscaled_confidence_sigma=conf, degrees_of_freedom=self.m - t)


def set_threads(n: int):
    lib().vo_set_threads(int(n))


# ---------------------------------------------------------------------------
# workloads of the reference's tests/benches (shared_test_code/src/lib.rs)
# ---------------------------------------------------------------------------
def ref_linspace(first, last, count):
    """shared_test_code/src/lib.rs:20-34 -- bug-compatible: first + (first-last)/(count-1)*n."""
    n = np.arange(count, dtype=np.float64)
    return first + (first - last) / (count - 1) * n


DOUBLE_EXP_OFFSET = [(EXP_DECAY, [0]), (EXP_DECAY, [1]), (CONSTANT, [])]          # lib.rs:119-135
DOUBLE_EXP_OFFSET_TEST_HELPER = [(EXP_DECAY, [1]), (EXP_DECAY, [0]), (CONSTANT, [])]  # src/test_helpers/mod.rs:56-71
OLEARY = [(EXP_RATE_COS, [1, 2]), (EXP_RATE_COS, [0, 1])]                         # models.rs:321-322
