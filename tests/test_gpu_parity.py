"""GPU parity tests: the CUDA path (through the C ABI) against the CPU oracle on identical
seeded inputs, and against the reference's golden vectors.

Tolerances (BASELINE.json north_star): nonlinear parameters within 1e-8 relative, residual
norm within 1e-10 relative to ||Y_w||; element-wise quantities (residuals, Jacobian,
coefficients) to 1e-9 relative to their scale. fp32 problems use stated looser bounds.
"""
import os
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


@pytest.fixture
def ctx_options():
    """Set tunables of the default context for one test (vp_ctx_set_option) and restore the defaults."""
    import varpro_b200 as vb
    defaults = {"fit_mode": "auto", "eval_kernel": "fused", "stream_kernel": "auto", "panel_generic": 0, "max_ctas": 0}
    touched = []

    def set_option(key, value, slot=0):
        touched.append((key, slot))
        vb.set_option(key, value, slot=slot)
    yield set_option
    for key, slot in touched:
        vb.set_option(key, defaults[key], slot=slot)


def test_generic_kernel_matches_fast_kernel(ctx_options):
    wl = W.c2(S=40)
    fast = W.make_gpu_problem(wl)
    ctx_options("stream_kernel", "generic")
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


def test_fused_kernel_matches_split_kernels(ctx_options):
    """fit_kernel_dmma (panel fused into the streaming pass) against K1 + K2 launched separately."""
    wl = W.c2(S=200)
    fused = W.make_gpu_problem(wl)
    ctx_options("eval_kernel", "split")
    split = W.make_gpu_problem(wl)
    ctx_options("eval_kernel", "fused")
    for alpha in ([2.0, 6.5], [1.1, 2.9]):
        fused.set_params(alpha)
        split.set_params(alpha)
        rf, rs = fused.reduce(), split.reduce()
        assert abs(rf["rnorm2"] - rs["rnorm2"]) <= 1e-12 * rs["rnorm2"]
        assert np.max(np.abs(rf["H"] - rs["H"])) <= 1e-12 * np.abs(rs["H"]).max()
        assert np.max(np.abs(rf["g"] - rs["g"])) <= 1e-9 * np.abs(rs["g"]).max()
        Cf, Cs = fused.linear_coefficients(), split.linear_coefficients()
        assert np.max(np.abs(Cf - Cs)) <= 1e-12 * np.abs(Cs).max()


@pytest.mark.parametrize("mode", ["auto", "graph", "host"])
def test_fit_modes_agree(ctx_options, mode):
    """The persistent whole-fit kernel (auto), the CUDA-graph loop (the driver of model shapes without a fused
    kernel) and the host-driven loop run the same lmder state machine on the same reductions."""
    import varpro_b200 as vb
    wl = W.c2(S=96)
    ctx_options("fit_mode", mode)
    if mode != "auto":
        ctx_options("eval_kernel", "split")
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
    """vp_fit_many (work-queue kernel: all SMs serve all fits) against one vp_fit per problem."""
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


@pytest.mark.parametrize("items_per_cta", [1, 2, 5])
def test_fit_many_is_bitwise_vp_fit(ctx_options, items_per_cta):
    """The work-queue kernel folds the same canonical parts in the same order as the persistent fit kernel and runs
    the same LM code: identical parameters, coefficients, objective and evaluation counts, whatever the work-item
    size -- on the noise-free benchmark data too, where rounding decides the last LM steps."""
    import varpro_b200 as vb
    vb.set_option("queue_items_per_cta", items_per_cta)
    try:
        solver = vb.LevMarSolver.default()
        rng = np.random.default_rng(5)
        wls = [W.c2(S=1500 + 40 * k, seed=300 + k) for k in range(5)]
        for k, wl in enumerate(wls[1:], 1):  # different truths / noise => different LM paths and evaluation counts
            x = wl["x"]
            tau = (1.0 + 0.2 * k, 3.0 + 0.5 * k)
            Phi = np.stack([np.exp(-x / tau[0]), np.exp(-x / tau[1]), np.ones_like(x)], axis=1)
            wl["Y"] = np.asfortranarray(Phi @ wl["C_true"] + (1e-3 * k) * rng.standard_normal((x.shape[0], wl["Y"].shape[1])))
        seq = [solver.fit(W.make_gpu_problem(wl)) for wl in wls]
        many = solver.fit_many([W.make_gpu_problem(wl) for wl in wls])
        assert len({r.minimization_report.number_of_evaluations for r in seq}) > 1
        for a, b in zip(seq, many):
            assert a.was_successful() and b.was_successful()
            assert np.array_equal(a.nonlinear_parameters(), b.nonlinear_parameters())
            assert a.minimization_report.number_of_evaluations == b.minimization_report.number_of_evaluations
            assert a.minimization_report.objective_function == b.minimization_report.objective_function
            assert np.array_equal(a.linear_coefficients(), b.linear_coefficients())
    finally:
        vb.set_option("queue_items_per_cta", 2)


def test_fp32_fit_many_is_bitwise_vp_fit_on_the_persistent_fp32_kernel():
    """fp32 problems (BASELINE config 4 shape, weighted): vp_fit runs the persistent whole-fit kernel instantiated for
    float storage, vp_fit_many the work-queue kernel: same parts, same tile code => bitwise equal."""
    import varpro_b200 as vb
    solver = vb.LevMarSolver.default()
    wls = []
    for k in range(4):
        wl = W.c4(S=600 + 200 * k, seed=40 + k)
        wls.append(wl)
    seq = [solver.fit(W.make_gpu_problem(wl, dtype=np.float32)) for wl in wls]
    many = solver.fit_many([W.make_gpu_problem(wl, dtype=np.float32) for wl in wls])
    for a, b in zip(seq, many):
        assert a.was_successful() and b.was_successful()
        assert np.array_equal(a.nonlinear_parameters(), b.nonlinear_parameters())
        assert a.minimization_report.number_of_evaluations == b.minimization_report.number_of_evaluations
        assert np.array_equal(a.linear_coefficients(), b.linear_coefficients())


def test_problems_are_reusable_after_fit_many():
    """vp_fit_many leaves every problem ready for the next call: set_params back to the start gives bitwise the
    evaluation of a fresh problem (the per-problem ticket is re-armed), and a second fit_many repeats the first."""
    import varpro_b200 as vb
    solver = vb.LevMarSolver.default()
    wls = [W.c2(S=900 + 100 * k, seed=500 + k) for k in range(7)]
    probs = [W.make_gpu_problem(wl) for wl in wls]
    fresh = [W.make_gpu_problem(wl).reduce() for wl in wls]
    first = solver.fit_many(probs)
    ref = [(r.nonlinear_parameters(), r.minimization_report.number_of_evaluations) for r in first]
    for rep in range(3):
        for p, wl, fr in zip(probs, wls, fresh):
            p.set_params(wl["alpha0"])
            red = p.reduce()
            assert red["rnorm2"] == fr["rnorm2"] and np.array_equal(red["H"], fr["H"]) and np.array_equal(red["g"], fr["g"])
        again = solver.fit_many(probs)
        for (a, nf), b in zip(ref, again):
            assert np.array_equal(a, b.nonlinear_parameters()) and nf == b.minimization_report.number_of_evaluations


def test_fit_many_is_deterministic_and_bitwise_vp_fit_on_the_2048_row_tiling():
    """m = 1500 runs the 16-warp tiling of the fused kernels (1024 < m <= 2048): repeated vp_fit_many launches and
    vp_fit must agree bitwise there too (different thread count, same canonical partition)."""
    import varpro_b200 as vb
    solver = vb.LevMarSolver.default()
    rng = np.random.default_rng(9)
    m = 1500
    x = np.linspace(0.0, 12.0, m)
    wls = []
    for k in range(6):
        tau = (1.0 + 0.1 * k, 3.0 + 0.3 * k)
        S = 700 + 64 * k
        Cs = rng.uniform(0.0, 100.0, size=(3, S))
        Phi = np.stack([np.exp(-x / tau[0]), np.exp(-x / tau[1]), np.ones_like(x)], axis=1)
        wls.append(dict(x=x, Y=np.asfortranarray(Phi @ Cs), basis=W.DOUBLE_EXP, q=2, alpha0=[2.0, 6.5], weights=None))
    seq = [solver.fit(W.make_gpu_problem(wl)) for wl in wls]
    ref = [(r.nonlinear_parameters(), r.minimization_report.number_of_evaluations, r.linear_coefficients()) for r in seq]
    for rep in range(4):
        many = solver.fit_many([W.make_gpu_problem(wl) for wl in wls])
        for (a, nf, Cc), b in zip(ref, many):
            assert b.was_successful()
            assert np.array_equal(a, b.nonlinear_parameters()), rep
            assert nf == b.minimization_report.number_of_evaluations, rep
            assert np.array_equal(Cc, b.linear_coefficients()), rep


# fp32 (BASELINE config 4). The reference is generic over the scalar but no reference test runs an f32
# fit (SURVEY.md 8c "unpinned"), so fp32 results are pinned against the fp64 oracle on the SAME
# (fp32-rounded) inputs. Stated tolerances: the kernels keep Q, E and Y in fp32 (eps = 6e-8) and
# accumulate in fp64, so element-wise quantities agree to ~1e-5 of their scale and the converged
# parameters to 5e-4 relative (LM stops on ftol = xtol = 30*eps_f32 = 3.6e-6).
F32_PARAM_REL = 5e-4
F32_STATE_REL = 2e-5


def test_c4_fp32_weighted_state_and_fit_small_S():
    import varpro_b200 as vb
    wl = W.c4(S=48)
    gp = W.make_gpu_problem(wl, dtype=np.float32)
    wl64 = dict(wl, x=wl["x"].astype(np.float64), Y=np.asfortranarray(wl["Y"].astype(np.float64)),
                weights=wl["weights"].astype(np.float64))
    op = W.make_oracle(wl64)
    Yw = wl64["weights"][:, None] * wl64["Y"]
    r_g, r_o = gp.residuals().astype(np.float64), op.residuals()
    assert np.max(np.abs(r_g - r_o)) <= F32_STATE_REL * np.abs(Yw).max()
    C_g, C_o = gp.linear_coefficients().astype(np.float64), op.linear_coefficients()
    assert np.max(np.abs(C_g - C_o)) <= 1e-3 * np.abs(C_o).max()
    red = gp.reduce()
    J_o = op.jacobian()
    assert abs(red["rnorm2"] - r_o @ r_o) <= 1e-3 * (r_o @ r_o)
    assert np.max(np.abs(red["H"] - J_o.T @ J_o)) <= 1e-3 * np.abs(J_o.T @ J_o).max()
    res = vb.LevMarSolver.default().fit(gp)
    rep = op.fit()
    assert res.was_successful() and rep["successful"]
    a_g, a_o = res.nonlinear_parameters(), op.params()
    assert np.max(np.abs(a_g - a_o) / np.abs(a_o)) <= F32_PARAM_REL, (a_g, a_o)
    rn_g, rn_o = np.sqrt(2 * res.minimization_report.objective_function), np.sqrt(2 * rep["objective_function"])
    assert abs(rn_g - rn_o) <= 1e-4 * rn_o


def test_c4_fp32_full_size():
    """S = 16 384 in fp32 (65.5 MB): recovery of the generating parameters and coefficients; column 0 is
    the reference's lmfit asset."""
    import varpro_b200 as vb
    wl = W.c4()
    gp = W.make_gpu_problem(wl, dtype=np.float32)
    res = vb.LevMarSolver.default().fit(gp)
    assert res.was_successful()
    a = res.nonlinear_parameters()
    assert np.allclose(a, [2.4, 6.0], rtol=2e-3), a
    C = res.linear_coefficients().astype(np.float64)
    assert np.max(np.abs(C[:, 1:] - wl["C_gen"][:, 1:])) <= 0.05 * np.abs(wl["C_gen"]).max()
    chi2 = 2 * res.minimization_report.objective_function / (1000 * 16384 - 3 * 16384 - 2)
    assert 0.5 * 3.2e-5 < chi2 < 2 * 3.2e-5 + 1e-4  # weighted noise level of the generator (sigma = 0.01, w = 1/sqrt(y))
    # and directly against the fp64 oracle on the same fp32-rounded inputs (a few seconds with all host threads)
    from oracle import varpro_oracle as vo
    wl64 = dict(wl, x=wl["x"].astype(np.float64), Y=np.asfortranarray(wl["Y"].astype(np.float64)),
                weights=wl["weights"].astype(np.float64))
    vo.set_threads(os.cpu_count() or 1)
    op = W.make_oracle(wl64)
    rep = op.fit()
    vo.set_threads(1)
    assert rep["successful"]
    assert np.max(np.abs(a - op.params()) / np.abs(op.params())) <= F32_PARAM_REL
    rn_g, rn_o = np.sqrt(2 * res.minimization_report.objective_function), np.sqrt(2 * rep["objective_function"])
    assert abs(rn_g - rn_o) <= 1e-4 * rn_o
    assert np.max(np.abs(C - op.linear_coefficients())) <= 1e-3 * np.abs(op.linear_coefficients()).max()


def _batch_model(wl, m):
    import varpro_b200 as vb
    names = [f"p{k}" for k in range(wl["q"])]
    b = vb.SeparableModelBuilder(names)
    fns = {0: vb.ExpDecay, 1: vb.Constant, 2: vb.ExpRateCos, 3: vb.SinPhase}
    for kind, idx in wl["basis"]:
        b = b.function([names[i] for i in idx], fns[kind]()) if idx else b.invariant_function(fns[kind]())
    return b.independent_variable(wl["x"]).initial_parameters([1.0] * wl["q"]).build()


@pytest.mark.parametrize("m,P", [(256, 12), (1000, 5), (4096, 4)])  # 4096 = BASELINE config 3's sample count
def test_c3_independent_batch_matches_oracle_per_problem(m, P):
    """BASELINE config 3 shape at test size: every problem of the batch against its own oracle fit."""
    import varpro_b200 as vb
    wl = W.triple_exp_batch(P=P, m=m)
    batch = vb.IndependentBatch(_batch_model(wl, m), wl["Y"], wl["alpha0"])
    res = batch.fit()
    assert res.successful.all(), res.terminations
    for p in range(P):
        one = dict(x=wl["x"], Y=np.asfortranarray(wl["Y"][:, p:p + 1]), basis=wl["basis"], q=3,
                   alpha0=list(wl["alpha0"][p]), weights=None)
        op = W.make_oracle(one)
        rep = op.fit()
        assert rep["successful"]
        a_g, a_o = np.sort(res.nonlinear_parameters[p]), np.sort(op.params())
        # noisy data, ftol stop: see SURVEY 7.2-8. When two decay components merge, one tau runs off to
        # +-1e14 (exp(-x/tau) -> the constant 1): its value is not determined by the data -- both
        # implementations must agree that it ran away, and agree on the determined parameters.
        det = np.abs(a_o) < 1e6
        assert np.array_equal(det, np.abs(a_g) < 1e6), (p, a_g, a_o)
        assert np.max(np.abs(a_g[det] - a_o[det]) / np.abs(a_o[det])) <= 1e-7, (p, a_g, a_o)
        rn_g, rn_o = np.sqrt(2 * res.objective_function[p]), np.sqrt(2 * rep["objective_function"])
        assert abs(rn_g - rn_o) <= REL_RNORM * np.linalg.norm(one["Y"]), p
        if det.all():
            c_g = res.linear_coefficients[:, p][np.argsort(res.nonlinear_parameters[p])]
            c_o = op.linear_coefficients()[:, 0][np.argsort(op.params())]
            assert np.max(np.abs(c_g - c_o)) <= 1e-6 * np.abs(c_o).max()
    batch.close()


def test_c3_independent_batch_weighted_double_exp_equals_single_problem_path():
    """The batch kernel against the MRHS/single-problem kernels on the same weighted problems."""
    import varpro_b200 as vb
    rng = np.random.default_rng(11)
    m, P = 300, 9
    x = np.linspace(0.0, 10.0, m)
    w = rng.uniform(0.5, 1.5, size=m)
    tau = np.array([1.0, 3.0]) * rng.uniform(0.8, 1.25, size=(P, 2))
    Y = np.stack([4.0 * np.exp(-x / t[0]) + 2.5 * np.exp(-x / t[1]) + 1.0 for t in tau], axis=1)
    Y = np.asfortranarray(Y + 1e-3 * rng.standard_normal(Y.shape))
    wl = dict(x=x, basis=W.DOUBLE_EXP, q=2)
    a0 = tau * np.array([1.3, 0.8])
    batch = vb.IndependentBatch(_batch_model(wl, m), Y, a0, weights=w)
    res = batch.fit()
    assert res.successful.all()
    for p in range(P):
        one = dict(x=x, Y=np.asfortranarray(Y[:, p:p + 1]), basis=W.DOUBLE_EXP, q=2, alpha0=list(a0[p]), weights=w)
        r1 = vb.LevMarSolver.default().fit(W.make_gpu_problem(one))
        assert np.max(np.abs(res.nonlinear_parameters[p] - r1.nonlinear_parameters()) / r1.nonlinear_parameters()) <= 1e-7
        assert np.max(np.abs(res.linear_coefficients[:, p] - r1.linear_coefficients())) <= 1e-6
    batch.close()


@pytest.mark.parametrize("weighted", [False, True])
def test_statistics_lmfit_goldens(weighted):
    """fit_with_statistics against the reference's lmfit goldens (tests/integration_tests/main.rs:600-612,
    :670-687): chi2_red 1e-8, covariance 1e-6, 88% confidence band 1e-6 -- and against the oracle."""
    import varpro_b200 as vb
    wl = W.lmfit_case(weighted)
    gp = W.make_gpu_problem(wl)
    res, st = vb.LevMarSolver.default().fit_with_statistics(gp)
    assert abs(st.reduced_chi2() - wl["gold"]["chi2"]) <= 1e-8
    # lmfit orders (c1, c2, c3, tau1, tau2) == the reference's (c..., alpha...) ordering (main.rs:603)
    assert np.max(np.abs(st.covariance_matrix() - wl["covmat"])) <= 1e-6
    assert np.max(np.abs(st.confidence_band_radius(0.88) - wl["conf"])) <= 1e-6
    op = W.make_oracle(wl)
    op.fit()
    so = op.statistics(0)
    # both fits stop within the north-star tolerance of each other (parameters 1e-8 relative): the covariance at the
    # two stopping points agrees to that order, not to rounding
    assert np.max(np.abs(st.covariance_matrix() - so["covariance"])) <= 1e-7 * np.abs(so["covariance"]).max() + 1e-12
    assert abs(st.reduced_chi2() - so["reduced_chi2"]) <= 1e-9 * so["reduced_chi2"]


def test_statistics_oleary_goldens():
    import varpro_b200 as vb
    wl = W.oleary()
    gp = W.make_gpu_problem(wl)
    res, st = vb.LevMarSolver.default().fit_with_statistics(gp)
    assert abs(st.regression_standard_error() - wl["sigma"]) <= 1e-5                    # main.rs:780-784
    cov = st.covariance_matrix()
    assert np.allclose(cov, wl["cov"], rtol=0, atol=1e-5)                               # main.rs:786-802
    assert np.allclose(st.calculate_correlation_matrix(), wl["corr"], rtol=0, atol=1e-4)  # main.rs:804-823


def test_statistics_mrhs_every_column_matches_oracle():
    """The reference supports SingleRhs statistics only; per column with the shared alpha the GPU result
    must equal the oracle's try_calculate applied to that column (BASELINE config 4 semantics)."""
    import varpro_b200 as vb
    rng = np.random.default_rng(5)
    m, S = 200, 7
    x = np.linspace(0.0, 12.0, m)
    w = rng.uniform(0.5, 1.5, size=m)
    Cs = rng.uniform(1.0, 5.0, size=(3, S))
    Phi = np.stack([np.exp(-x / 1.5), np.exp(-x / 4.0), np.ones_like(x)], axis=1)
    Y = np.asfortranarray(Phi @ Cs + 1e-2 * rng.standard_normal((m, S)))
    wl = dict(x=x, Y=Y, basis=W.DOUBLE_EXP, q=2, alpha0=[1.2, 5.0], weights=w)
    gp, op = W.make_gpu_problem(wl), W.make_oracle(wl)
    res, sts = vb.LevMarSolver.default().fit_with_statistics(gp)
    op.fit()
    assert len(sts) == S
    for s in range(S):
        so = op.statistics(s)
        assert np.max(np.abs(sts[s].covariance_matrix() - so["covariance"])) <= 1e-6 * np.abs(so["covariance"]).max()
        assert abs(sts[s].reduced_chi2() - so["reduced_chi2"]) <= 1e-8 * so["reduced_chi2"]
        assert np.max(np.abs(sts[s]._sigma - so["unscaled_confidence_sigma"])) <= 1e-6 * np.abs(so["unscaled_confidence_sigma"]).max()


def _closure_double_exp_model(x, alpha0):
    """The reference's own way of building the model: closures + partial derivatives
    (shared_test_code/src/lib.rs:101-135)."""
    import varpro_b200 as vb

    def exp_decay(x, tau):
        return np.exp(-x / tau)

    def exp_decay_dtau(x, tau):
        return np.exp(-x / tau) * x / (tau * tau)

    return (vb.SeparableModelBuilder(["tau1", "tau2"])
            .function(["tau1"], exp_decay).partial_deriv("tau1", exp_decay_dtau)
            .function(["tau2"], exp_decay).partial_deriv("tau2", exp_decay_dtau)
            .invariant_function(lambda x: np.ones_like(x))
            .independent_variable(x).initial_parameters(alpha0).build())


def test_host_evaluated_closure_model_matches_builtin_and_oracle():
    """Arbitrary SeparableNonlinearModel implementations (host closures) through vp_model_create_hosteval:
    Phi / dPhi come from the host once per evaluation, the O(m*S) pass and the LM step stay in the library."""
    import varpro_b200 as vb
    wl = W.c2(S=48)
    model = _closure_double_exp_model(wl["x"], wl["alpha0"])
    assert model.is_host_evaluated()
    gp = vb.SeparableProblemBuilder.mrhs(model).observations(wl["Y"]).build()
    op = W.make_oracle(wl)
    _compare_state(gp, op, np.linalg.norm(wl["Y"]), "closure model @alpha0")
    builtin = W.make_gpu_problem(wl)
    rb, rh = builtin.reduce(), gp.reduce()
    assert abs(rb["rnorm2"] - rh["rnorm2"]) <= 1e-10 * rb["rnorm2"]
    assert np.max(np.abs(rb["H"] - rh["H"])) <= 1e-9 * np.abs(rb["H"]).max()
    res = vb.LevMarSolver.default().fit(gp)
    assert res.was_successful()
    assert np.allclose(np.sort(res.nonlinear_parameters()), [1.0, 3.0], rtol=0, atol=1e-8)
    assert np.max(np.abs(res.best_fit() - wl["Y"])) <= 1e-5 * np.abs(wl["Y"]).max()
    rep = op.fit()
    a_g, a_o = np.sort(res.nonlinear_parameters()), np.sort(op.params())
    assert np.max(np.abs(a_g - a_o) / a_o) <= REL_PARAM


def test_host_evaluated_model_statistics_and_errors():
    import varpro_b200 as vb
    wl = W.lmfit_case(True)
    model = _closure_double_exp_model(wl["x"], wl["alpha0"])
    gp = vb.SeparableProblemBuilder.new(model).observations(wl["Y"][:, 0]).weights(wl["weights"]).build()
    res, st = vb.LevMarSolver.default().fit_with_statistics(gp)
    assert np.allclose(res.nonlinear_parameters(), wl["gold"]["tau"], rtol=0, atol=1e-5)
    assert np.max(np.abs(st.covariance_matrix() - wl["covmat"])) <= 1e-6
    # a closure that raises: the cache becomes None, fit stops with a User termination (levmar/mod.rs:43-45)
    calls = {"n": 0}

    def flaky(x, tau):
        calls["n"] += 1
        if calls["n"] > 3:
            raise RuntimeError("model error")
        return np.exp(-x / tau)

    bad = (vb.SeparableModelBuilder(["tau"]).function(["tau"], flaky)
           .partial_deriv("tau", lambda x, tau: np.exp(-x / tau) * x / tau ** 2)
           .invariant_function(lambda x: np.ones_like(x))
           .independent_variable(wl["x"]).initial_parameters([1.0]).build())
    gp2 = vb.SeparableProblemBuilder.new(bad).observations(wl["Y"][:, 0]).build()
    with pytest.raises(vb.FitError) as ei:
        vb.LevMarSolver.default().fit(gp2)
    assert repr(ei.value.result.minimization_report.termination).lower().find("user") >= 0
    assert gp2.residuals() is None
    # missing derivative -> build error like the reference
    with pytest.raises(vb.ModelBuildError):
        (vb.SeparableModelBuilder(["tau"]).function(["tau"], lambda x, tau: np.exp(-x / tau))
         .independent_variable(wl["x"]).initial_parameters([1.0]).build())


def test_device_resident_materialisers_match_host_copies():
    """vp_residuals_device / vp_jacobian_device / vp_best_fit_device write into a caller-owned device buffer."""
    import ctypes as C
    import torch
    from varpro_b200 import _lib
    wl = W.c2(S=72)
    gp = W.make_gpu_problem(wl)
    gp.set_params([1.4, 3.3])
    m, S, q = 1024, 72, 2
    lib = _lib.load()
    for fn, host, count in ((lib.vp_residuals_device, gp.residuals(), m * S),
                            (lib.vp_jacobian_device, gp.jacobian().ravel(order="F"), m * S * q),
                            (lib.vp_best_fit_device, gp.best_fit().ravel(order="F"), m * S)):
        buf = torch.empty(count, dtype=torch.float64, device="cuda")
        assert fn(gp._h, C.c_void_p(buf.data_ptr())) == 0
        torch.cuda.synchronize()
        assert np.array_equal(buf.cpu().numpy(), np.asarray(host).ravel(order="F"))


def test_fp32_problems_share_the_work_queue_kernel():
    """fp32 problems in vp_fit_many: streamed as fp32, all arithmetic in fp64 (work-queue kernel)."""
    import varpro_b200 as vb
    solver = vb.LevMarSolver.default()
    wls = [W.c4(S=40 + 8 * k, seed=50 + k) for k in range(4)]
    many = solver.fit_many([W.make_gpu_problem(wl, dtype=np.float32) for wl in wls])
    for wl, r in zip(wls, many):
        assert r.was_successful()
        wl64 = dict(wl, x=wl["x"].astype(np.float64), Y=np.asfortranarray(wl["Y"].astype(np.float64)),
                    weights=wl["weights"].astype(np.float64))
        op = W.make_oracle(wl64)
        assert op.fit()["successful"]
        assert np.max(np.abs(r.nonlinear_parameters() - op.params()) / np.abs(op.params())) <= F32_PARAM_REL


def test_problems_of_different_sizes_share_kernel_instantiations():
    """Regression (found by scripts/stress_fit_many.py): the dynamic shared-memory limit of a kernel is a
    process-wide attribute. A problem that needs less shared memory, created later, must not lower the
    limit under a problem planned with more (m = 500 and m = 200 use the same instantiation)."""
    import varpro_b200 as vb
    rng = np.random.default_rng(3)
    solver = vb.LevMarSolver.default()

    def make(m, S):
        x = np.linspace(0.0, 10.0, m)
        Phi = np.stack([np.exp(-x / 1.0), np.exp(-x / 3.0), np.ones_like(x)], axis=1)
        Y = np.asfortranarray(Phi @ rng.uniform(1.0, 5.0, size=(3, S)) + 1e-3 * rng.standard_normal((m, S)))
        return W.make_gpu_problem(dict(x=x, Y=Y, basis=W.DOUBLE_EXP, q=2, alpha0=[1.3, 2.4], weights=None))

    big = [make(500, 50) for _ in range(3)]
    small = [make(200, 50) for _ in range(3)]       # planned later, smaller shared-memory footprint
    for p in big + small:
        r = solver.fit(p)
        assert r.was_successful() and np.allclose(np.sort(r.nonlinear_parameters()), [1.0, 3.0], atol=1e-3)
    big2 = [make(500, 60) for _ in range(3)]
    small2 = [make(200, 60) for _ in range(3)]
    res = solver.fit_many(big2) + solver.fit_many(small2) + solver.fit_many([make(500, 40), make(500, 44)])
    assert all(r.was_successful() for r in res)


def _noisy_weighted_problem(m=160, S=5, noise=0.3, seed=21):
    rng = np.random.default_rng(seed)
    x = np.linspace(0.0, 10.0, m)
    Phi = np.stack([np.exp(-x / 1.2), np.exp(-x / 3.7), np.ones_like(x)], axis=1)
    Y = np.asfortranarray(Phi @ rng.uniform(1.0, 5.0, size=(3, S)) + noise * rng.standard_normal((m, S)))
    w = rng.uniform(0.5, 1.5, size=m)
    return dict(x=x, Y=Y, basis=W.DOUBLE_EXP, q=2, alpha0=[0.9, 4.5], weights=w)


def test_full_golub_pereyra_jacobian_is_the_exact_derivative_of_the_residual():
    """VP_JACOBIAN_FULL (the term the reference leaves as a TODO, src/solvers/levmar/mod.rs:188-190): the
    explicit Jacobian must equal central finite differences of residuals() AWAY from the optimum on a
    large-residual problem -- which Kaufman's approximation (the default, = the reference) does not --
    and (||r||^2, J^T r, J^T J) must be those of that Jacobian."""
    wl = _noisy_weighted_problem()
    gp = W.make_gpu_problem(wl)
    alpha = np.array([1.0, 4.2])

    def resid(a):
        gp.set_params(a)
        return gp.residuals().copy()

    J_fd = np.empty((160 * 5, 2))
    for k in range(2):
        h = 1e-6 * alpha[k]
        ap, am = alpha.copy(), alpha.copy()
        ap[k] += h
        am[k] -= h
        J_fd[:, k] = (resid(ap) - resid(am)) / (2 * h)
    gp.set_params(alpha)
    J_kaufman = gp.jacobian()
    gp.set_jacobian("full")
    J_full = gp.jacobian()
    scale = np.abs(J_fd).max()
    assert np.max(np.abs(J_full - J_fd)) <= 1e-6 * scale
    assert np.max(np.abs(J_kaufman - J_fd)) >= 1e-3 * scale      # the approximation really differs here
    r = gp.residuals()
    red = gp.reduce()
    assert np.max(np.abs(red["H"] - J_full.T @ J_full)) <= 1e-9 * np.abs(J_full.T @ J_full).max()
    assert np.max(np.abs(red["g"] - J_full.T @ r)) <= 1e-8 * np.abs(J_kaufman.T @ r).max() + 1e-9
    gp.set_jacobian("kaufman")
    red_k = gp.reduce()
    assert np.max(np.abs(red_k["H"] - J_kaufman.T @ J_kaufman)) <= 1e-9 * np.abs(red_k["H"]).max()


def test_full_jacobian_fit_reaches_the_same_minimum():
    import varpro_b200 as vb
    wl = _noisy_weighted_problem(m=300, S=8, noise=0.5, seed=4)
    a = vb.LevMarSolver.default().fit(W.make_gpu_problem(wl))
    gp = W.make_gpu_problem(wl).set_jacobian("full")
    b = vb.LevMarSolver.default().fit(gp)
    assert a.was_successful() and b.was_successful()
    pa, pb = a.nonlinear_parameters(), b.nonlinear_parameters()
    assert np.max(np.abs(pa - pb) / np.abs(pa)) <= 1e-6
    assert abs(a.minimization_report.objective_function - b.minimization_report.objective_function) \
        <= 1e-10 * a.minimization_report.objective_function
    # the exact Jacobian converges quadratically near the solution: no more evaluations than Kaufman's
    assert b.minimization_report.number_of_evaluations <= a.minimization_report.number_of_evaluations + 1


def test_c2_full_size_matches_the_oracle():
    """BASELINE config 2 at FULL size (m = 1024, S = 4096): the oracle needs about a second, so the fit is
    compared directly -- parameters to 1e-8 relative, residual norm to 1e-10 ||Y||, every coefficient."""
    import varpro_b200 as vb
    from oracle import varpro_oracle as vo
    wl = W.c2(S=4096)
    vo.set_threads(os.cpu_count() or 1)
    op = W.make_oracle(wl)
    rep = op.fit()
    vo.set_threads(1)
    assert rep["successful"]
    single = vb.LevMarSolver.default().fit(W.make_gpu_problem(wl))                      # persistent kernel
    many = vb.LevMarSolver.default().fit_many([W.make_gpu_problem(wl) for _ in range(3)])  # work-queue kernel
    Yn = np.linalg.norm(wl["Y"])
    a_o = np.sort(op.params())
    C_o = op.linear_coefficients()
    if op.params()[0] > op.params()[1]:
        C_o = C_o[[1, 0, 2]]
    for r in [single] + many:
        assert r.was_successful()
        a = r.nonlinear_parameters()
        C = r.linear_coefficients()
        if a[0] > a[1]:
            a, C = a[::-1], C[[1, 0, 2]]
        assert np.max(np.abs(a - a_o) / a_o) <= REL_PARAM
        rn = np.sqrt(2 * r.minimization_report.objective_function)
        assert abs(rn - np.sqrt(2 * rep["objective_function"])) <= REL_RNORM * Yn
        assert np.max(np.abs(C - C_o)) <= 1e-7 * np.abs(C_o).max()
