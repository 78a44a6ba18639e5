"""GPU parity tests: the CUDA path (through the C ABI) against the CPU oracle on identical
seeded inputs, and against the reference's golden vectors.

Tolerances (BASELINE.json north_star): nonlinear parameters within 1e-8 relative, residual
norm within 1e-10 relative to ||Y_w||; element-wise quantities (residuals, Jacobian,
coefficients) to 1e-9 relative to their scale. fp32 problems use stated looser bounds.
"""
import numpy as np
import pytest

import workloads as W

pytestmark = pytest.mark.gpu

REL_PARAM = 1e-8
REL_RNORM = 1e-10


def _sorted_pair(alpha, C):
    """tau labels may switch (tests/integration_tests/main.rs:135-141)."""
    if alpha[0] > alpha[1]:
        return alpha[::-1].copy(), C[[1, 0, 2]]
    return alpha, C


def _compare_state(gp, op, Yw_norm, tag):
    """residuals / jacobian / coefficients / reduction of the GPU problem vs the oracle at the same alpha."""
    r_g, r_o = gp.residuals(), op.residuals()
    assert r_g is not None and r_o is not None
    assert np.max(np.abs(r_g - r_o)) <= 1e-9 * max(1.0, Yw_norm), tag
    assert abs(np.linalg.norm(r_g) - np.linalg.norm(r_o)) <= REL_RNORM * Yw_norm, tag
    J_g, J_o = gp.jacobian(), op.jacobian()
    scale = max(1.0, np.abs(J_o).max())
    assert np.max(np.abs(J_g - J_o)) <= 1e-9 * scale, tag
    C_g, C_o = gp.linear_coefficients(), op.linear_coefficients()
    C_g = C_g.reshape(C_o.shape)
    assert np.max(np.abs(C_g - C_o)) <= 1e-8 * max(1.0, np.abs(C_o).max()), tag
    red = gp.reduce()
    g_o = J_o.T @ r_o
    H_o = J_o.T @ J_o
    assert abs(red["rnorm2"] - r_o @ r_o) <= 1e-9 * max(r_o @ r_o, (REL_RNORM * Yw_norm) ** 2), tag
    assert np.max(np.abs(red["H"] - H_o)) <= 1e-9 * np.abs(H_o).max(), tag
    assert np.max(np.abs(red["g"] - g_o)) <= 1e-9 * (np.linalg.norm(J_o, axis=0).max() * max(np.linalg.norm(r_o), 1e-300)) \
        + 1e-12 * np.abs(J_o).max() * Yw_norm, tag


@pytest.mark.parametrize("weighted", [False, True])
def test_octave_residual_goldens(weighted):
    wl = W.octave_case(weighted)
    gp = W.make_gpu_problem(wl)
    gp.set_params(wl["alpha_eval"])
    r = gp.residuals()
    assert np.max(np.abs(r - wl["expected_residuals"])) <= wl["tol"]
    op = W.make_oracle(wl)
    op.set_params(wl["alpha_eval"])
    Yw = wl["Y"][:, 0] * (wl["weights"] if weighted else 1.0)
    _compare_state(gp, op, np.linalg.norm(Yw), "octave")


@pytest.mark.parametrize("S", [2, 3])
def test_state_parity_mrhs20_both_jacobian_branches(S):
    wl = W.mrhs20(S)
    gp, op = W.make_gpu_problem(wl), W.make_oracle(wl)
    for alpha in ([2.5, 6.5], [0.7, 4.1], [1.0, 3.0]):
        gp.set_params(alpha)
        op.set_params(alpha)
        _compare_state(gp, op, np.linalg.norm(wl["Y"]), f"mrhs20 S={S} alpha={alpha}")


def test_c1_fit_matches_reference_goldens_and_oracle():
    wl = W.c1()
    import varpro_b200 as vb
    gp = W.make_gpu_problem(wl)
    res = vb.LevMarSolver.default().fit(gp)
    assert res.was_successful()
    alpha, C = _sorted_pair(res.nonlinear_parameters(), res.linear_coefficients())
    # reference test: tests/integration_tests/main.rs:152-156 (epsilon = 1e-8)
    assert np.allclose(alpha, wl["alpha_true"], rtol=0, atol=1e-8)
    assert np.allclose(C, wl["C_true"][:, 0], rtol=0, atol=1e-8)
    assert np.max(np.abs(res.best_fit() - wl["Y"][:, 0])) <= 1e-5 * 1.0 + 1e-9 * np.abs(wl["Y"]).max()
    op = W.make_oracle(wl)
    rep = op.fit()
    assert rep["successful"]
    a_o, _ = _sorted_pair(op.params(), op.linear_coefficients()[:, 0])
    assert np.max(np.abs(alpha - a_o) / np.abs(a_o)) <= REL_PARAM
    rn_g = np.sqrt(2 * res.minimization_report.objective_function)
    rn_o = np.sqrt(2 * rep["objective_function"])
    assert abs(rn_g - rn_o) <= REL_RNORM * np.linalg.norm(wl["Y"])


@pytest.mark.parametrize("S", [2, 3])
def test_mrhs20_fit(S):
    import varpro_b200 as vb
    wl = W.mrhs20(S)
    gp = W.make_gpu_problem(wl)
    res = vb.LevMarSolver.default().fit(gp)
    alpha = res.nonlinear_parameters()
    C = res.linear_coefficients()
    if alpha[0] > alpha[1]:
        alpha, C = alpha[::-1], C[[1, 0, 2]]
    assert np.allclose(alpha, [1.0, 3.0], rtol=0, atol=1e-8)      # main.rs:455-462 / :540-550
    assert np.allclose(C, wl["C_true"], rtol=0, atol=1e-8)
    assert np.max(np.abs(res.best_fit() - wl["Y"])) <= 1e-5


@pytest.mark.parametrize("weighted", [False, True])
def test_lmfit_goldens(weighted):
    import varpro_b200 as vb
    wl = W.lmfit_case(weighted)
    gp = W.make_gpu_problem(wl)
    res = vb.LevMarSolver.default().fit(gp)
    tau, c = res.nonlinear_parameters(), res.linear_coefficients()
    assert np.allclose(tau, wl["gold"]["tau"], rtol=0, atol=1e-5)   # main.rs:594-598 / :664-668
    assert np.allclose(c, wl["gold"]["c"], rtol=0, atol=1e-5)
    op = W.make_oracle(wl)
    op.fit()
    assert np.max(np.abs(tau - op.params()) / np.abs(op.params())) <= REL_PARAM
    Yw = wl["Y"][:, 0] * (wl["weights"] if weighted else 1.0)
    rn_g = np.sqrt(2 * res.minimization_report.objective_function)
    assert abs(rn_g - np.linalg.norm(op.residuals())) <= REL_RNORM * np.linalg.norm(Yw)
    chi2 = 2 * res.minimization_report.objective_function / (1000 - 5)
    assert abs(chi2 - wl["gold"]["chi2"]) <= 1e-8                  # main.rs:600 / :670


def test_oleary_goldens():
    import varpro_b200 as vb
    wl = W.oleary()
    gp = W.make_gpu_problem(wl)
    res = vb.LevMarSolver.default().fit(gp)
    assert np.allclose(res.nonlinear_parameters(), wl["alpha_true"], rtol=0, atol=1e-5)  # main.rs:754
    assert np.allclose(res.linear_coefficients(), wl["c_true"], rtol=0, atol=1e-5)       # main.rs:755
    assert np.allclose(gp.residuals(), wl["wresid"], rtol=0, atol=1e-5)                   # main.rs:769-778
    assert np.max(np.abs(res.best_fit() - wl["Y"][:, 0])) <= 1e-2                         # main.rs:743
    op = W.make_oracle(wl)
    op.fit()
    assert np.max(np.abs(res.nonlinear_parameters() - op.params()) / np.abs(op.params())) <= REL_PARAM


def test_c2_shape_small_S_fit_and_state():
    """C2 generator at S=64 (oracle finishes in seconds): state parity at the start, fit parity at the end."""
    import varpro_b200 as vb
    wl = W.c2(S=64)
    gp, op = W.make_gpu_problem(wl), W.make_oracle(wl)
    _compare_state(gp, op, np.linalg.norm(wl["Y"]), "c2 S=64 @alpha0")
    res = vb.LevMarSolver.default().fit(gp)
    rep = op.fit()
    assert res.was_successful() and rep["successful"]
    a_g, a_o = np.sort(res.nonlinear_parameters()), np.sort(op.params())
    assert np.max(np.abs(a_g - a_o) / a_o) <= REL_PARAM
    assert np.allclose(a_g, [1.0, 3.0], rtol=0, atol=1e-8)
    rn_g = np.sqrt(2 * res.minimization_report.objective_function)
    rn_o = np.sqrt(2 * rep["objective_function"])
    assert abs(rn_g - rn_o) <= REL_RNORM * np.linalg.norm(wl["Y"])


def test_c2_full_size_properties():
    """BASELINE config 2 at full size (S=4096): size-independent properties instead of the oracle."""
    import varpro_b200 as vb
    wl = W.c2(S=4096)
    gp = W.make_gpu_problem(wl)
    res = vb.LevMarSolver.default().fit(gp)
    assert res.was_successful()
    alpha = res.nonlinear_parameters()
    C = res.linear_coefficients()
    if alpha[0] > alpha[1]:
        alpha, C = alpha[::-1], C[[1, 0, 2]]
    assert np.allclose(alpha, [1.0, 3.0], rtol=0, atol=1e-8)
    assert np.max(np.abs(C - wl["C_true"])) <= 1e-6            # max|C-C*| ~ 2e-9
    # the residual is the projection: ||R||_F / ||Y||_F at machine-precision level
    rn = np.sqrt(2 * res.minimization_report.objective_function)
    assert rn <= 1e-13 * np.linalg.norm(wl["Y"])
    # linearity of the inner solve: coefficients of (Y1 + 2*Y2) = C1 + 2*C2 at fixed alpha
    Y = wl["Y"]
    a = [1.3, 2.7]
    g1 = W.make_gpu_problem(wl, alpha0=a, Y=np.asfortranarray(Y[:, :128]))
    g2 = W.make_gpu_problem(wl, alpha0=a, Y=np.asfortranarray(Y[:, 128:256]))
    g3 = W.make_gpu_problem(wl, alpha0=a, Y=np.asfortranarray(Y[:, :128] + 2.0 * Y[:, 128:256]))
    C1, C2, C3 = g1.linear_coefficients(), g2.linear_coefficients(), g3.linear_coefficients()
    assert np.max(np.abs(C3 - (C1 + 2 * C2))) <= 1e-9 * np.abs(C3).max()
    # idempotence: the residual of the residual is the residual (P_perp^2 = P_perp)
    R1 = g1.residuals().reshape(128, 1024).T
    gr = W.make_gpu_problem(wl, alpha0=a, Y=np.asfortranarray(R1))
    R2 = gr.residuals().reshape(128, 1024).T
    assert np.max(np.abs(R2 - R1)) <= 1e-9 * np.abs(Y[:, :128]).max()


def test_generic_kernel_matches_fast_kernel(monkeypatch):
    wl = W.c2(S=40)
    fast = W.make_gpu_problem(wl)
    monkeypatch.setenv("VP_STREAM_GENERIC", "1")
    gen = W.make_gpu_problem(wl)
    rf, rg = fast.reduce(), gen.reduce()
    assert abs(rf["rnorm2"] - rg["rnorm2"]) <= 1e-10 * rf["rnorm2"]
    assert np.max(np.abs(rf["H"] - rg["H"])) <= 1e-10 * np.abs(rf["H"]).max()
    assert np.max(np.abs(rf["g"] - rg["g"])) <= 1e-9 * np.abs(rf["g"]).max()
    assert np.max(np.abs(fast.linear_coefficients() - gen.linear_coefficients())) <= 1e-9 * 100


def test_ragged_and_edge_shapes():
    """odd m (padding row), S not a multiple of the tile width, S=1, tiny m."""
    rng = np.random.default_rng(7)
    for m, S in [(21, 1), (33, 7), (255, 13), (1000, 9), (1024, 37)]:
        x = np.linspace(0.1, 9.0, m)
        Cs = rng.uniform(0.5, 2.0, size=(3, S))
        Phi = np.stack([np.exp(-x / 1.5), np.exp(-x / 4.0), np.ones_like(x)], axis=1)
        Y = Phi @ Cs + 1e-3 * rng.standard_normal((m, S))
        w = rng.uniform(0.5, 1.5, size=m)
        wl = dict(x=x, Y=np.asfortranarray(Y), basis=W.DOUBLE_EXP, q=2, alpha0=[1.2, 5.0], weights=w)
        gp, op = W.make_gpu_problem(wl), W.make_oracle(wl)
        _compare_state(gp, op, np.linalg.norm(w[:, None] * Y), f"ragged m={m} S={S}")


def test_nonfinite_parameters_give_none_and_numerical_termination():
    """tau = 0 makes exp(-x/tau) non-finite: cache is None (src/solvers/levmar/mod.rs:43-45,70-72)."""
    import varpro_b200 as vb
    wl = W.mrhs20(2)
    gp = W.make_gpu_problem(wl)
    gp.set_params([0.0, 3.0])
    assert gp.residuals() is None and gp.jacobian() is None and gp.linear_coefficients() is None
    with pytest.raises(vb.FitError) as ei:
        vb.LevMarSolver.default().fit(gp)
    assert not ei.value.result.was_successful()


def test_fused_kernel_matches_split_kernels(monkeypatch):
    """fit_kernel_dmma (panel fused into the streaming pass) against K1 + K2 launched separately."""
    wl = W.c2(S=200)
    fused = W.make_gpu_problem(wl)
    monkeypatch.setenv("VP_EVAL_KERNEL", "split")
    split = W.make_gpu_problem(wl)
    monkeypatch.delenv("VP_EVAL_KERNEL")
    for alpha in ([2.0, 6.5], [1.1, 2.9]):
        fused.set_params(alpha)
        split.set_params(alpha)
        rf, rs = fused.reduce(), split.reduce()
        assert abs(rf["rnorm2"] - rs["rnorm2"]) <= 1e-12 * rs["rnorm2"]
        assert np.max(np.abs(rf["H"] - rs["H"])) <= 1e-12 * np.abs(rs["H"]).max()
        assert np.max(np.abs(rf["g"] - rs["g"])) <= 1e-9 * np.abs(rs["g"]).max()
        Cf, Cs = fused.linear_coefficients(), split.linear_coefficients()
        assert np.max(np.abs(Cf - Cs)) <= 1e-12 * np.abs(Cs).max()


@pytest.mark.parametrize("mode", ["persistent", "graph", "host"])
def test_fit_modes_agree(monkeypatch, mode):
    """The persistent whole-fit kernel, the CUDA-graph loop and the host-driven loop run the same
    lmder state machine on the same reductions."""
    import varpro_b200 as vb
    wl = W.c2(S=96)
    monkeypatch.setenv("VP_FIT_MODE", mode)
    if mode != "persistent":
        monkeypatch.setenv("VP_EVAL_KERNEL", "split")
    gp = W.make_gpu_problem(wl)
    res = vb.LevMarSolver.default().fit(gp)
    assert res.was_successful()
    assert np.allclose(np.sort(res.nonlinear_parameters()), [1.0, 3.0], rtol=0, atol=1e-8)
    C = res.linear_coefficients()
    if res.nonlinear_parameters()[0] > res.nonlinear_parameters()[1]:
        C = C[[1, 0, 2]]
    assert np.max(np.abs(C - wl["C_true"])) <= 1e-6


def test_sharded_fit_world1_exercises_the_mailbox_protocol():
    """A communicator of world size 1: every evaluation goes through the NVLink mailbox exchange
    (with itself) inside the kernel; results must equal the plain problem bit for bit."""
    import varpro_b200 as vb
    from varpro_b200 import sharding
    wl = W.c2(S=128)
    plain = W.make_gpu_problem(wl)
    comm = sharding.Communicator(0, 1)
    sh = comm.attach(W.make_gpu_problem(wl))
    rp, rs = plain.reduce(), sh.reduce()
    assert rp["rnorm2"] == rs["rnorm2"] and np.array_equal(rp["H"], rs["H"]) and np.array_equal(rp["g"], rs["g"])
    a = vb.LevMarSolver.default().fit(plain)
    b = vb.LevMarSolver.default().fit(sh)
    assert np.array_equal(a.nonlinear_parameters(), b.nonlinear_parameters())
    assert a.minimization_report.number_of_evaluations == b.minimization_report.number_of_evaluations
    sh.close()
    comm.close()


def test_fit_many_equals_sequential_fits():
    """vp_fit_many (concurrent persistent kernels on SM slices) against one vp_fit per problem."""
    import varpro_b200 as vb
    solver = vb.LevMarSolver.default()
    wls = [W.c2(S=64 + 8 * k, seed=100 + k) for k in range(6)] + [W.mrhs20(3), W.lmfit_case(True)]
    seq = [solver.fit(W.make_gpu_problem(wl)) for wl in wls]
    many = solver.fit_many([W.make_gpu_problem(wl) for wl in wls])
    for wl, a, b in zip(wls, seq, many):
        assert b.was_successful()
        pa, pb = np.sort(a.nonlinear_parameters()), np.sort(b.nonlinear_parameters())
        assert np.max(np.abs(pa - pb) / np.abs(pa)) <= REL_PARAM
        Yn = np.linalg.norm(wl["Y"])
        rn_a = np.sqrt(2 * a.minimization_report.objective_function)
        rn_b = np.sqrt(2 * b.minimization_report.objective_function)
        assert abs(rn_a - rn_b) <= REL_RNORM * Yn
    # more problems than SMs: the surplus queues behind the running fits
    wl = W.mrhs20(2)
    res = solver.fit_many([W.make_gpu_problem(wl) for _ in range(200)])
    assert all(r.was_successful() for r in res)
    assert all(np.allclose(np.sort(r.nonlinear_parameters()), [1.0, 3.0], atol=1e-8) for r in res)
