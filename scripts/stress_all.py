"""Broader randomized checks on the GPU box (beyond tests/): independent-batch kernel vs the single-problem
path over random shapes, statistics vs the oracle, host-evaluated models vs built-in kinds, world-1
communicator, several host threads fitting concurrently on their own contexts."""
import os
import sys
import time
from concurrent.futures import ThreadPoolExecutor

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import workloads as W  # noqa: E402
import varpro_b200 as vb  # noqa: E402
from varpro_b200 import sharding  # noqa: E402

rng = np.random.default_rng(int(sys.argv[1]) if len(sys.argv) > 1 else 5)
solver = vb.LevMarSolver.default()
t0 = time.time()


def dexp_problem(m, S, noise=1e-3, weights=False):
    x = np.linspace(0.0, 10.0, m)
    tau = np.array([1.0, 3.0]) * rng.uniform(0.85, 1.2, size=2)
    Phi = np.stack([np.exp(-x / tau[0]), np.exp(-x / tau[1]), np.ones_like(x)], axis=1)
    Y = np.asfortranarray(Phi @ rng.uniform(1.0, 5.0, size=(3, S)) + noise * rng.standard_normal((m, S)))
    w = rng.uniform(0.5, 1.5, size=m) if weights else None
    return dict(x=x, Y=Y, basis=W.DOUBLE_EXP, q=2, alpha0=list(tau * rng.uniform(0.8, 1.3, size=2)), weights=w)


# 1. independent batch vs single-problem path, random m (row tilings 512/1024/2048/4096) and P
for m in (37, 300, 700, 1500, 3000):
    P = int(rng.integers(3, 40))
    wl = dexp_problem(m, P, weights=bool(rng.integers(0, 2)))
    names = ["p0", "p1"]
    model = (vb.SeparableModelBuilder(names).function(["p0"], vb.ExpDecay()).function(["p1"], vb.ExpDecay())
             .invariant_function(vb.Constant()).independent_variable(wl["x"]).initial_parameters([1.0, 3.0]).build())
    a0 = np.tile(np.array(wl["alpha0"]), (P, 1)) * rng.uniform(0.95, 1.05, size=(P, 2))
    batch = vb.IndependentBatch(model, wl["Y"], a0, weights=wl["weights"])
    res = batch.fit()
    for p in range(P):
        one = dict(wl, Y=np.asfortranarray(wl["Y"][:, p:p + 1]), alpha0=list(a0[p]))
        try:
            r1 = solver.fit(W.make_gpu_problem(one))
        except vb.FitError as e:
            r1 = e.result
        assert bool(res.successful[p]) == r1.was_successful(), (m, p)
        if r1.was_successful():
            rel = np.max(np.abs(res.nonlinear_parameters[p] - r1.nonlinear_parameters()) / np.abs(r1.nonlinear_parameters()))
            assert rel <= 1e-6, (m, p, rel)
    batch.close()
print(f"batch vs single ok ({time.time()-t0:.1f} s)")

# 2. statistics vs oracle on random weighted MRHS problems
for m, S in ((50, 3), (333, 9), (1000, 5)):
    wl = dexp_problem(m, S, noise=1e-2, weights=True)
    gp, op = W.make_gpu_problem(wl), W.make_oracle(wl)
    _, sts = solver.fit_with_statistics(gp)
    op.fit()
    for s in range(S):
        so = op.statistics(s)
        assert np.max(np.abs(sts[s].covariance_matrix() - so["covariance"])) <= 1e-6 * np.abs(so["covariance"]).max(), (m, s)
print(f"statistics vs oracle ok ({time.time()-t0:.1f} s)")

# 3. host-evaluated closure model vs built-in kinds
for m, S in ((64, 5), (777, 33)):
    wl = dexp_problem(m, S)
    model = (vb.SeparableModelBuilder(["a", "b"])
             .function(["a"], lambda x, t: np.exp(-x / t)).partial_deriv("a", lambda x, t: np.exp(-x / t) * x / (t * t))
             .function(["b"], vb.ExpDecay())
             .invariant_function(vb.Constant())
             .independent_variable(wl["x"]).initial_parameters(wl["alpha0"]).build())
    rh = solver.fit(vb.SeparableProblemBuilder.mrhs(model).observations(wl["Y"]).build())
    rb = solver.fit(W.make_gpu_problem(wl))
    assert np.max(np.abs(rh.nonlinear_parameters() - rb.nonlinear_parameters()) / rb.nonlinear_parameters()) <= 1e-7
print(f"host-evaluated vs built-in ok ({time.time()-t0:.1f} s)")

# 4. world-1 communicator on random sizes
comm = sharding.Communicator(0, 1)
for m, S in ((100, 17), (1024, 300)):
    wl = dexp_problem(m, S)
    a = solver.fit(W.make_gpu_problem(wl))
    b = solver.fit(comm.attach(W.make_gpu_problem(wl)))
    assert np.array_equal(a.nonlinear_parameters(), b.nonlinear_parameters())
comm.close()
print(f"world-1 communicator ok ({time.time()-t0:.1f} s)")

# 5. several host threads, each with its own library context, fitting concurrently
wls = [dexp_problem(int(rng.choice([200, 500, 1024])), int(rng.integers(20, 400))) for _ in range(24)]
ref = [np.sort(solver.fit(W.make_gpu_problem(wl)).nonlinear_parameters()) for wl in wls]


def work(i):
    r = vb.LevMarSolver.default().fit(W.make_gpu_problem(wls[i], ctx_slot=1 + i % 4))
    return np.sort(r.nonlinear_parameters())


with ThreadPoolExecutor(4) as ex:
    got = list(ex.map(work, range(len(wls))))
for a, b in zip(ref, got):
    assert np.max(np.abs(a - b) / np.abs(a)) <= 1e-9
print(f"concurrent host threads ok ({time.time()-t0:.1f} s)")
print("stress_all ok")
