"""The CPU oracle (oracle/varpro_oracle.c) against every golden vector the reference's own tests
hold for the hot path. This is what pins the oracle."""
import numpy as np
import pytest
from scipy import stats

import workloads as W


def _sorted(alpha, c):
    if alpha[0] > alpha[1]:
        return alpha[::-1].copy(), c[[1, 0, 2]]
    return alpha, c


@pytest.mark.parametrize("weighted", [False, True])
def test_octave_residuals(weighted):
    """src/solvers/levmar/test.rs:145-162 (1e-4) and :177-207 (1e-3)."""
    wl = W.octave_case(weighted)
    op = W.make_oracle(wl)
    op.set_params(wl["alpha_eval"])
    assert np.max(np.abs(op.residuals() - wl["expected_residuals"])) <= wl["tol"]


def test_residuals_near_zero_at_true_parameters():
    """src/solvers/levmar/test.rs:133-142: residual ~0 at (2,4) for the 4-digit Octave data."""
    wl = W.octave_case(False)
    op = W.make_oracle(wl, alpha0=[2.0, 4.0])
    r = op.residuals()
    assert np.max(np.abs(r)) <= 1e-4 and r @ r <= 1e-8


def test_gradient_of_rss_matches_finite_difference_weighted():
    """src/solvers/levmar/test.rs:51-108: d||r||^2/dalpha_k = 2 r^T J_k vs a 6-point central difference."""
    wl = W.octave_case(True)
    wl = dict(wl, weights=np.sqrt(wl["Y"][:, 0]) + np.sin(wl["Y"][:, 0]))
    op = W.make_oracle(wl, alpha0=[1.0, 2.0])
    fixed = np.array([0.5, 7.5])
    h = np.sqrt(np.finfo(np.float64).eps)

    def rss(a):
        op.set_params(a)
        r = op.residuals()
        return r @ r

    def nd(k):
        def f(t):
            a = fixed.copy()
            a[k] = t
            return rss(a)
        p0 = fixed[k]
        return (-f(p0 - 3 * h) + 9 * f(p0 - 2 * h) - 45 * f(p0 - h) + 45 * f(p0 + h) - 9 * f(p0 + 2 * h) + f(p0 + 3 * h)) / (60 * h)

    num = [nd(0), nd(1)]
    op.set_params(fixed)
    r, J = op.residuals(), op.jacobian()
    ana = 2 * J.T @ r
    assert np.allclose(num, ana, rtol=1e-6, atol=1e-6)


def test_jacobian_matches_forward_difference_at_true_parameters():
    """src/solvers/levmar/test.rs:21-40 (Kaufman J equals the true one where the residual vanishes)."""
    wl = W.octave_case(False)
    op = W.make_oracle(wl, alpha0=[2.0, 4.0])
    J = op.jacobian()
    Jn = np.zeros_like(J)
    for k in range(2):
        a = np.array([2.0, 4.0])
        hk = 1e-7
        a[k] += hk
        op.set_params(a)
        rp = op.residuals()
        op.set_params([2.0, 4.0])
        Jn[:, k] = (rp - op.residuals()) / hk
    assert np.max(np.abs(J - Jn)) <= 1e-4  # forward difference of 4-digit data: the reference uses 1e-6 with its differentiator


def test_model_eval_column_order_and_values():
    """src/model/test.rs:105-127,176-246 (column order follows the order of .function calls)."""
    x = np.linspace(0.0, 3.0, 7)
    wl = dict(x=x, Y=np.ones((7, 1)), basis=W.DOUBLE_EXP_HELPER, q=2, alpha0=[1.5, 2.5], weights=None)
    op = W.make_oracle(wl)
    Phi = op.model_eval()
    # libm exp vs numpy's SIMD exp may differ by one ulp
    assert np.allclose(Phi[:, 0], np.exp(-x / 2.5), rtol=4e-16, atol=0) and np.allclose(Phi[:, 1], np.exp(-x / 1.5), rtol=4e-16, atol=0)
    assert np.array_equal(Phi[:, 2], np.ones(7))
    D0 = op.model_eval_partial_deriv(0)  # d/dtau1: only column 1 (tau1's function) is non-zero
    assert np.all(D0[:, 0] == 0) and np.all(D0[:, 2] == 0)
    assert np.allclose(D0[:, 1], np.exp(-x / 1.5) * x / 1.5 ** 2, rtol=1e-15, atol=0)


def test_c1_noise_free_fit():
    """tests/integration_tests/main.rs:93-157: tau=(1,3), c=(4,2.5,1) to 1e-8; best_fit to 1e-5."""
    wl = W.c1()
    op = W.make_oracle(wl)
    rep = op.fit()
    assert rep["successful"]
    a, c = _sorted(op.params(), op.linear_coefficients()[:, 0])
    assert np.allclose(a, [1.0, 3.0], rtol=0, atol=1e-8)
    assert np.allclose(c, [4.0, 2.5, 1.0], rtol=0, atol=1e-8)
    assert np.max(np.abs(op.best_fit()[:, 0] - wl["Y"][:, 0])) <= 1e-5


@pytest.mark.parametrize("S", [2, 3])
def test_mrhs20_fit_both_jacobian_branches(S):
    """tests/integration_tests/main.rs:399-463 (S=2 <= q) and :467-551 (S=3 > q)."""
    wl = W.mrhs20(S)
    op = W.make_oracle(wl)
    assert op.fit()["successful"]
    a, C = op.params(), op.linear_coefficients()
    if a[0] > a[1]:
        a, C = a[::-1], C[[1, 0, 2]]
    assert np.allclose(a, [1.0, 3.0], rtol=0, atol=1e-8)
    assert np.allclose(C, wl["C_true"], rtol=0, atol=1e-8)


@pytest.mark.parametrize("weighted", [False, True])
def test_lmfit_goldens_fit_and_statistics(weighted):
    """tests/integration_tests/main.rs:554-613 / :616-688."""
    wl = W.lmfit_case(weighted)
    op = W.make_oracle(wl)
    assert op.fit()["successful"]
    assert np.allclose(op.params(), wl["gold"]["tau"], rtol=0, atol=1e-5)
    assert np.allclose(op.linear_coefficients()[:, 0], wl["gold"]["c"], rtol=0, atol=1e-5)
    st = op.statistics(0)
    assert abs(st["reduced_chi2"] - wl["gold"]["chi2"]) <= 1e-8
    assert np.max(np.abs(st["covariance"] - wl["covmat"])) <= 1e-6
    band = stats.t.ppf((0.88 + 1) / 2, st["degrees_of_freedom"]) * st["unscaled_confidence_sigma"]
    assert np.max(np.abs(band - wl["conf"])) <= 1e-6


def test_oleary_goldens():
    """tests/integration_tests/main.rs:713-824."""
    wl = W.oleary()
    op = W.make_oracle(wl)
    assert op.fit()["successful"]
    assert np.allclose(op.params(), wl["alpha_true"], rtol=0, atol=1e-5)
    assert np.allclose(op.linear_coefficients()[:, 0], wl["c_true"], rtol=0, atol=1e-5)
    st = op.statistics(0)
    assert np.allclose(st["weighted_residuals"], wl["wresid"], rtol=0, atol=1e-5)
    assert np.allclose(op.residuals(), st["weighted_residuals"], rtol=0, atol=1e-5)
    assert abs(np.sqrt(st["reduced_chi2"]) - wl["sigma"]) <= 1e-5
    cov = st["covariance"]
    assert np.allclose(cov, wl["cov"], rtol=0, atol=1e-5)
    d = np.sqrt(np.diag(cov))
    assert np.allclose(cov / np.outer(d, d), wl["corr"], rtol=0, atol=1e-4)
    assert np.max(np.abs(op.best_fit()[:, 0] - wl["Y"][:, 0])) <= 1e-2


def test_against_scipy_minpack_lmder():
    """Independent cross-check of the LM restatement: scipy.optimize.leastsq drives the original
    Fortran lmder with the crate's defaults."""
    from scipy.optimize import leastsq
    eps = np.finfo(np.float64).eps
    for wl in (W.c1(), W.mrhs20(3), W.lmfit_case(False), W.lmfit_case(True), W.oleary(), W.c2(S=16)):
        op = W.make_oracle(wl)
        sp = W.make_oracle(wl)

        def f(a):
            sp.set_params(a)
            return sp.residuals().copy()

        def Df(a):
            sp.set_params(a)
            return np.array(sp.jacobian())

        q = wl["q"]
        xs, _, info, _, ier = leastsq(f, np.array(wl["alpha0"], float), Dfun=Df, full_output=True, ftol=30 * eps,
                                      xtol=30 * eps, gtol=30 * eps, factor=100.0, maxfev=100 * (q + 1))
        rep = op.fit()
        assert rep["successful"] and ier in (1, 2, 3, 4)
        assert np.max(np.abs(np.sort(op.params()) - np.sort(xs)) / np.abs(np.sort(xs))) <= 1e-8
        assert abs(rep["number_of_evaluations"] - info["nfev"]) <= 3


def test_builder_semantics():
    """src/problem/builder/test.rs:53-54,94-99: Y_w = W*y; epsilon = |epsilon|; invalid sizes give no problem."""
    from oracle import varpro_oracle as vo
    wl = W.octave_case(True)
    op = W.make_oracle(wl)
    op.set_params([1e9, 2e9])  # Phi ~ all ones: rank deficient, still returns a cache
    assert op.residuals() is not None
    with pytest.raises(ValueError):
        vo.OracleProblem(np.zeros(0), W.DOUBLE_EXP, 2, np.zeros((0, 1)), [1.0, 2.0])
