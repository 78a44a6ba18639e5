"""rank_policy.cuh (the inner solve's singular-value truncation) compiled for the host: against numpy's SVD.

Reference rule: sigma_i <= eps truncated in the solve (src/solvers/levmar/mod.rs:52-54); MATLAB rule: sigma_i <=
m * eps_machine * sigma_1 (matlab/varpro.m:642-643), encoded as a negative tolerance."""
import numpy as np

import lm_harness as LH


def _r_of(Phi):
    return np.triu(np.linalg.qr(Phi, mode="r"))


def test_full_rank_panels_take_the_cheap_path():
    rng = np.random.default_rng(0)
    for n in (1, 2, 3, 4, 6, 8):
        R = _r_of(rng.standard_normal((50, n)))
        full, trunc, _, _ = LH.rank_policy(R, np.finfo(float).eps)
        assert full and not trunc
        full, trunc, _, _ = LH.rank_policy(R, -50 * np.finfo(float).eps)
        assert full and not trunc


def test_truncated_solve_is_the_pseudo_inverse_of_the_reference_rule():
    """Two (nearly) equal columns: the truncated solve V Sigma^+ U^T equals numpy's pinv with the same threshold."""
    rng = np.random.default_rng(1)
    x = np.linspace(0, 10, 200)
    for delta, tol in [(0.0, 1e-8), (1e-11, 1e-8), (1e-14, -200 * np.finfo(float).eps), (0.0, -200 * np.finfo(float).eps)]:
        Phi = np.stack([np.exp(-x / 2.0), np.exp(-x / (2.0 + delta)), np.ones_like(x)], axis=1)
        R = _r_of(Phi)
        full, trunc, Urot, RinvEff = LH.rank_policy(R, tol)
        assert not full and trunc, (delta, tol)
        s = np.linalg.svd(R, compute_uv=False)
        thr = tol if tol >= 0 else -tol * s[0]
        keep = s > thr
        assert keep.sum() == 2
        # c = RinvEff (Urot^T b) must equal pinv_thr(R) b for every b
        P = RinvEff @ Urot.T
        Ur, sv, Vt = np.linalg.svd(R)
        P_ref = (Vt.T * np.where(sv > thr, 1.0 / sv, 0.0)) @ Ur.T
        assert np.max(np.abs(P - P_ref)) <= 1e-9 * np.abs(P_ref).max()
        # Urot has orthonormal kept columns and zero truncated ones
        G = Urot.T @ Urot
        assert np.allclose(np.sort(np.diag(G)), [0.0, 1.0, 1.0], atol=1e-12)
        b = rng.standard_normal(3)
        y_coef = P @ b
        assert abs(y_coef[0] - y_coef[1]) <= 1e-6 * max(1.0, abs(y_coef[0]))  # minimum norm: the twin columns share the load


def test_exact_zero_diagonal_is_truncated_under_either_rule():
    R = np.array([[2.0, 1.0, 0.5], [0.0, 0.0, 0.3], [0.0, 0.0, 1.5]])
    for tol in (np.finfo(float).eps, -100 * np.finfo(float).eps):
        full, trunc, Urot, RinvEff = LH.rank_policy(R, tol)
        assert not full and trunc
        s = np.linalg.svd(R, compute_uv=False)
        assert np.isfinite(RinvEff).all() and np.isfinite(Urot).all()
        assert np.linalg.matrix_rank(Urot) == int((s > 1e-12).sum())
