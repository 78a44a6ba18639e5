"""Workloads and golden constants transcribed from the reference's own tests/benches.

Each entry cites the reference file:line it restates. Used by the oracle tests (CPU)
and by the GPU parity tests, so both run on identical inputs.
"""
import os

import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def ref_linspace(first, last, count):
    """shared_test_code/src/lib.rs:20-34 -- bug-compatible (first + (first-last)/(count-1)*n):
    linspace(0., 12.5, 1024) is the DESCENDING grid 0 ... -12.5 (SURVEY.md gotcha G1)."""
    n = np.arange(count, dtype=np.float64)
    return first + (first - last) / (count - 1) * n


# basis tables: (kind, [parameter indices]) -- kinds: 0 ExpDecay, 1 Constant, 2 ExpRateCos, 3 SinPhase
DOUBLE_EXP = [(0, [0]), (0, [1]), (1, [])]            # shared_test_code/src/lib.rs:119-135 (tau1, tau2, 1)
DOUBLE_EXP_HELPER = [(0, [1]), (0, [0]), (1, [])]     # src/test_helpers/mod.rs:56-71 (tau2, tau1, 1)
OLEARY = [(2, [1, 2]), (2, [0, 1])]                   # shared_test_code/src/models.rs:321-322
TRIPLE_EXP = [(0, [0]), (0, [1]), (0, [2])]           # BASELINE config 3 (extension of exp_decay)


def double_exp_data(x, tau, c):
    return c[0] * np.exp(-x / tau[0]) + c[1] * np.exp(-x / tau[1]) + c[2]


def c1():
    """tests/integration_tests/main.rs:93-157 / benches/double_exponential_without_noise.rs:98-118."""
    x = ref_linspace(0.0, 12.5, 1024)
    y = double_exp_data(x, (1.0, 3.0), (4.0, 2.5, 1.0))
    return dict(x=x, Y=y[:, None], basis=DOUBLE_EXP, q=2, alpha0=[2.0, 6.5], weights=None,
                alpha_true=[1.0, 3.0], C_true=np.array([[4.0], [2.5], [1.0]]))


def mrhs20(S):
    """tests/integration_tests/main.rs:399-463 (S=2) and :467-551 (S=3 > q: the other Jacobian branch)."""
    x = ref_linspace(0.0, 12.5, 20)
    coeffs = {2: [(2.0, 4.0, 0.2), (5.0, 1.0, 9.0)],
              3: [(2.0, 4.0, 0.2), (10.0, 12.0, 18.0), (5.0, 1.0, 9.0)]}[S]
    Y = np.stack([double_exp_data(x, (1.0, 3.0), c) for c in coeffs], axis=1)
    return dict(x=x, Y=Y, basis=DOUBLE_EXP, q=2, alpha0=[2.5, 6.5], weights=None,
                alpha_true=[1.0, 3.0], C_true=np.array(coeffs).T)


def c2(S=4096, seed=2314093240213841123 % (2 ** 63), dtype=np.float64):
    """benches/multiple_right_hand_sides.rs:57-103 with S right-hand sides (BASELINE config 2).
    The bench draws C* with rand 0.8 StdRng, which is not reproducible without Rust (gotcha G2);
    numpy PCG64 with the same literal seed (mod 2^63) is used instead."""
    x = ref_linspace(0.0, 12.5, 1024)
    rng = np.random.Generator(np.random.PCG64(seed))
    Cs = rng.uniform(0.0, 100.0, size=(3, S))
    Phi = np.stack([np.exp(-x / 1.0), np.exp(-x / 3.0), np.ones_like(x)], axis=1)
    Y = np.asfortranarray(Phi @ Cs)
    return dict(x=x.astype(dtype), Y=Y.astype(dtype), basis=DOUBLE_EXP, q=2, alpha0=[2.0, 6.5], weights=None,
                alpha_true=[1.0, 3.0], C_true=Cs)


def octave_case(weighted):
    """src/solvers/levmar/test.rs:112-162 (unweighted) and :164-207 (weighted)."""
    t = np.arange(11.0)
    y = np.array([4.0000, 2.9919, 2.3423, 1.9186, 1.6386, 1.4507, 1.3227, 1.2342, 1.1720, 1.1276, 1.0956])
    if weighted:
        w = np.sqrt(y) + 2.0 * np.sin(y)
        expected = np.array([-0.307187, 0.493658, 0.286886, -0.150538, -0.346541, -0.342850, -0.235283,
                             -0.084548, 0.077943, 0.237072, 0.385972])
        tol = 1e-3
    else:
        w = None
        expected = np.array([-0.032243, 0.236772, 0.028277, -0.105709, -0.149393, -0.136205, -0.092002,
                             -0.032946, 0.031394, 0.095542, 0.156511])
        tol = 1e-4
    return dict(x=t, Y=y[:, None], basis=DOUBLE_EXP_HELPER, q=2, alpha0=[2.0, 4.0], weights=w,
                alpha_eval=[0.5, 6.5], expected_residuals=expected, tol=tol)


def lmfit_case(weighted):
    """tests/integration_tests/main.rs:554-613 (unweighted) and :616-688 (weights 1/sqrt(y))."""
    name = "lmfit_weighted_multiexp_decay.npz" if weighted else "lmfit_multiexp_decay.npz"
    d = np.load(os.path.join(GOLDEN, name))
    x, y = d["x"], d["y"]
    if weighted:
        gold = dict(c=[2.24275841, 6.75609070, 1.59790007], tau=[2.43119160, 6.02052311], chi2=3.2117e-5)
        w = 1.0 / np.sqrt(y)
    else:
        gold = dict(c=[2.19344628, 6.80462652, 1.59995673], tau=[2.40392137, 5.99571068], chi2=1.0109e-4)
        w = None
    return dict(x=x, Y=y[:, None], basis=DOUBLE_EXP, q=2, alpha0=[1.0, 7.0], weights=w, gold=gold,
                covmat=d["covmat"], conf=d["conf"])


def oleary():
    """tests/integration_tests/main.rs:713-824; data from matlab/examples/varpro_example.m:26-57."""
    t = np.array([0.0, 0.1, 0.22, 0.31, 0.46, 0.50, 0.63, 0.78, 0.85, 0.97])
    y = np.array([6.9842, 5.1851, 2.8907, 1.4199, -0.2473, -0.5243, -1.0156, -1.0260, -0.9165, -0.6805])
    w = np.array([1.0, 1.0, 1.0, 0.5, 0.5, 1.0, 0.5, 1.0, 0.5, 0.5])
    return dict(
        x=t, Y=y[:, None], basis=OLEARY, q=3, alpha0=[0.5, 2.0, 3.0], weights=w,
        alpha_true=np.array([1.0132255e+00, 2.4968675e+00, 4.0625148e+00]),
        c_true=np.array([5.8416357e+00, 1.1436854e+00]),
        wresid=np.array([-1.1211e-03, 3.1751e-03, -2.7656e-03, -1.4600e-03, 1.2081e-03, 2.2586e-03,
                         -1.1101e-03, -2.2554e-03, 1.3257e-03, 1.4716e-03]),
        sigma=2.7539e-03,
        cov=np.array([[4.4887e-03, -4.4309e-03, -2.1613e-04, -4.6980e-04, -1.9052e-03],
                      [-4.4309e-03, 4.3803e-03, 2.1087e-04, 4.7170e-04, 1.8828e-03],
                      [-2.1613e-04, 2.1087e-04, 2.6925e-04, -3.6450e-05, 5.1919e-05],
                      [-4.6980e-04, 4.7170e-04, -3.6450e-05, 8.5784e-05, 2.0534e-04],
                      [-1.9052e-03, 1.8828e-03, 5.1919e-05, 2.0534e-04, 8.2272e-04]]),
        corr=np.array([[1.0000, -0.9993, -0.1966, -0.7571, -0.9914],
                       [-0.9993, 1.0000, 0.1942, 0.7695, 0.9918],
                       [-0.1966, 0.1942, 1.0000, -0.2398, 0.1103],
                       [-0.7571, 0.7695, -0.2398, 1.0000, 0.7729],
                       [-0.9914, 0.9918, 0.1103, 0.7729, 1.0000]]))


def triple_exp_batch(P=8, m=256, seed=65536):
    """BASELINE config 3 shape at test size: independent triple-exponential problems (not in the
    reference; SURVEY.md 8d C3 generator)."""
    rng = np.random.Generator(np.random.PCG64(seed))
    x = np.linspace(0.0, 20.0, m)
    tau = np.array([1.0, 3.0, 9.0]) * rng.uniform(0.8, 1.25, size=(P, 3))
    c = rng.uniform(1.0, 10.0, size=(P, 3))
    Y = np.stack([sum(c[p, j] * np.exp(-x / tau[p, j]) for j in range(3)) for p in range(P)], axis=1)
    Y = Y + rng.normal(0.0, 1e-3, size=Y.shape)
    return dict(x=x, Y=np.asfortranarray(Y), basis=TRIPLE_EXP, q=3, tau_true=tau, c_true=c,
                alpha0=tau * np.array([1.3, 0.8, 1.2]))


# ---------------------------------------------------------------------------
# adapters
# ---------------------------------------------------------------------------
def make_oracle(wl, alpha0=None, Y=None):
    from oracle import varpro_oracle as vo
    return vo.OracleProblem(wl["x"], wl["basis"], wl["q"], wl["Y"] if Y is None else Y,
                            wl["alpha0"] if alpha0 is None else alpha0, weights=wl["weights"])


_KIND_TO_FN = None


def make_gpu_problem(wl, alpha0=None, Y=None, dtype=np.float64, device=0, ctx_slot=0):
    """Build the same problem through the product's reference-facing API."""
    import varpro_b200 as vb
    fns = {0: vb.ExpDecay, 1: vb.Constant, 2: vb.ExpRateCos, 3: vb.SinPhase}
    names = [f"p{k}" for k in range(wl["q"])]
    b = vb.SeparableModelBuilder(names, dtype=dtype)
    for kind, idx in wl["basis"]:
        if idx:
            b = b.function([names[i] for i in idx], fns[kind]())
        else:
            b = b.invariant_function(fns[kind]())
    model = (b.independent_variable(np.asarray(wl["x"], dtype=dtype))
             .initial_parameters(list(wl["alpha0"] if alpha0 is None else alpha0)).build())
    Yv = wl["Y"] if Y is None else Y
    single = Yv.shape[1] == 1
    pb = vb.SeparableProblemBuilder.new(model) if single else vb.SeparableProblemBuilder.mrhs(model)
    pb = pb.observations(Yv[:, 0] if single else Yv)
    if wl.get("weights") is not None:
        pb = pb.weights(wl["weights"])
    return pb.device(device, ctx_slot).build()


def c4(S=16384, seed=16384, dtype=np.float32):
    """BASELINE config 4: x, y0 = the reference's weighted lmfit asset
    (test_assets/weighted_multiexp_decay), w = 1/sqrt(y0) shared by all right-hand sides; column 0 = y0
    (pinned by the lmfit goldens), columns s >= 1 = Phi(2.4, 6.0) c_s + N(0, 0.01^2) with
    c_s = (2.2, 6.8, 1.6) * U[0.5, 1.5)^3. Global fit from alpha0 = (1, 7); cast to `dtype`."""
    d = np.load(os.path.join(GOLDEN, "lmfit_weighted_multiexp_decay.npz"))
    x, y0 = d["x"], d["y"]
    rng = np.random.Generator(np.random.PCG64(seed))
    Phi = np.stack([np.exp(-x / 2.4), np.exp(-x / 6.0), np.ones_like(x)], axis=1)
    Cs = np.array([2.2, 6.8, 1.6])[:, None] * rng.uniform(0.5, 1.5, size=(3, S))
    Y = Phi @ Cs + rng.normal(0.0, 0.01, size=(x.shape[0], S))
    Y[:, 0] = y0
    w = 1.0 / np.sqrt(y0)
    return dict(x=x.astype(dtype), Y=np.asfortranarray(Y.astype(dtype)), basis=DOUBLE_EXP, q=2, alpha0=[1.0, 7.0],
                weights=w.astype(dtype), C_gen=Cs)
