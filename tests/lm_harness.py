"""ctypes access to tests/csrc/lm_harness.cpp (the product's LM state machine compiled for the host)."""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "csrc", "lm_harness.cpp")
LIB = os.path.join(HERE, "csrc", "liblm_harness.so")
DEP = os.path.join(HERE, "..", "varpro_b200", "csrc", "lm_step.cuh")
DEP2 = os.path.join(HERE, "..", "varpro_b200", "csrc", "rank_policy.cuh")


def build():
    if (not os.path.exists(LIB)) or os.path.getmtime(LIB) < max(os.path.getmtime(SRC), os.path.getmtime(DEP), os.path.getmtime(DEP2)):
        subprocess.run(["/usr/bin/g++", "-O2", "-std=c++17", "-x", "c++", "-shared", "-fPIC", "-o", LIB, SRC],
                       check=True, capture_output=True)
    return LIB


_lib = None


def lib():
    global _lib
    if _lib is None:
        L = C.CDLL(build())
        dp = C.POINTER(C.c_double)
        L.lmh_new.restype = C.c_void_p
        L.lmh_new.argtypes = [C.c_int, dp, C.c_double, C.c_double, C.c_double, C.c_double, C.c_int, C.c_int, C.c_double]
        L.lmh_free.argtypes = [C.c_void_p]
        L.lmh_advance.argtypes = [C.c_void_p, C.c_double, dp, dp, C.c_int]
        for n in ("lmh_trial", "lmh_accepted"):
            getattr(L, n).argtypes = [C.c_void_p, dp]
        for n in ("lmh_termination", "lmh_nfev", "lmh_last_accepted"):
            getattr(L, n).argtypes = [C.c_void_p]
        for n in ("lmh_fnorm", "lmh_par", "lmh_delta"):
            getattr(L, n).argtypes = [C.c_void_p]
            getattr(L, n).restype = C.c_double
        L.rph_policy.restype = C.c_int
        L.rph_policy.argtypes = [C.c_int, dp, dp, C.c_double, C.POINTER(C.c_int), dp, dp]
        L.lmh_set_generic.argtypes = [C.c_void_p, C.c_int]
        L.lmh_state_bytes.restype = C.c_int
        L.lmh_state.argtypes = [C.c_void_p, C.c_void_p]
        _lib = L
    return _lib


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


class LmHarness:
    def __init__(self, x0, ftol=None, xtol=None, gtol=None, stepbound=100.0, patience=100, scale_diag=True, generic=False):
        eps = np.finfo(np.float64).eps
        x0 = np.ascontiguousarray(x0, dtype=np.float64)
        self.q = len(x0)
        self._h = lib().lmh_new(self.q, _dp(x0), ftol or 30 * eps, xtol or 30 * eps, gtol or 30 * eps, stepbound,
                                patience * (self.q + 1), int(scale_diag), eps)
        if generic:
            lib().lmh_set_generic(self._h, 1)

    def __del__(self):
        if getattr(self, "_h", None):
            lib().lmh_free(self._h)

    def advance(self, rnorm2, g, H, finite=True):
        g = np.ascontiguousarray(g, dtype=np.float64)
        Hf = np.asfortranarray(H, dtype=np.float64)
        return bool(lib().lmh_advance(self._h, float(rnorm2), _dp(g), _dp(Hf), int(finite)))

    def trial(self):
        x = np.empty(self.q)
        lib().lmh_trial(self._h, _dp(x))
        return x

    def accepted(self):
        x = np.empty(self.q)
        lib().lmh_accepted(self._h, _dp(x))
        return x

    def state_bytes(self):
        """The raw LmState (bitwise comparison of the q-specialised and the generic instantiation)."""
        buf = C.create_string_buffer(lib().lmh_state_bytes())
        lib().lmh_state(self._h, buf)
        return buf.raw

    termination = property(lambda self: lib().lmh_termination(self._h))
    nfev = property(lambda self: lib().lmh_nfev(self._h))
    last_accepted = property(lambda self: bool(lib().lmh_last_accepted(self._h)))
    fnorm = property(lambda self: lib().lmh_fnorm(self._h))
    par = property(lambda self: lib().lmh_par(self._h))
    delta = property(lambda self: lib().lmh_delta(self._h))


def fit_with_oracle_evals(op, x0, **kw):
    """Run the product's LM state machine with evaluations from the oracle's explicit r and J."""
    h = LmHarness(x0, **kw)
    trace = []
    more = True
    x = np.asarray(x0, dtype=np.float64)
    while more:
        ok = op.set_params(x)
        r, J = op.residuals(), op.jacobian()
        if not ok or r is None:
            more = h.advance(np.nan, np.zeros(h.q), np.zeros((h.q, h.q)), finite=False)
        else:
            more = h.advance(r @ r, J.T @ r, J.T @ J, finite=True)
        trace.append(dict(x=x.copy(), fnorm=float(np.linalg.norm(r)) if r is not None else np.nan, par=h.par,
                          delta=h.delta, accepted=h.last_accepted))
        x = h.trial()
    return h, trace


def rank_policy(R, tol):
    """rank_policy.cuh on the host: (surely_full, truncated, Urot, RinvEff) for an upper triangular R."""
    R = np.asfortranarray(R, dtype=np.float64)
    n = R.shape[0]
    with np.errstate(divide="ignore", invalid="ignore"):
        Rinv = np.asfortranarray(np.linalg.inv(R) if np.all(np.diag(R) != 0) else np.full((n, n), np.inf))
    trunc = C.c_int(0)
    U = np.zeros((n, n), order="F")
    Ri = np.zeros((n, n), order="F")
    full = lib().rph_policy(n, _dp(R), _dp(Rinv), float(tol), C.byref(trunc), _dp(U), _dp(Ri))
    return bool(full), bool(trunc.value), U, Ri
