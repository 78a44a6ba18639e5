"""Stress test of the work-queue kernel: random batches of problems (random K, S, noise, fp64/fp32 mix) fitted
with vp_fit_many and compared with one vp_fit per problem. Run on the GPU box."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import workloads as W  # noqa: E402
import varpro_b200 as vb  # noqa: E402

rng = np.random.default_rng(int(sys.argv[1]) if len(sys.argv) > 1 else 1)
rounds = int(sys.argv[2]) if len(sys.argv) > 2 else 40
solver = vb.LevMarSolver.default()
t0 = time.time()
worst = 0.0
nfits = 0
for it in range(rounds):
    K = int(rng.integers(2, 41))
    m = int(rng.choice([64, 200, 500, 1000, 1024]))
    dtype = np.float32 if rng.random() < 0.25 else np.float64
    wls = []
    for k in range(K):
        S = int(rng.integers(1, 600))
        x = np.linspace(0.0, 10.0, m)
        tau = np.array([1.0, 3.0]) * rng.uniform(0.8, 1.25, size=2)
        Cs = rng.uniform(1.0, 5.0, size=(3, S))
        Phi = np.stack([np.exp(-x / tau[0]), np.exp(-x / tau[1]), np.ones_like(x)], axis=1)
        Y = np.asfortranarray(Phi @ Cs + 1e-3 * rng.standard_normal((m, S)))
        wls.append(dict(x=x.astype(dtype), Y=Y.astype(dtype), basis=W.DOUBLE_EXP, q=2,
                        alpha0=list(tau * rng.uniform(0.7, 1.4, size=2)), weights=None))
    many = solver.fit_many([W.make_gpu_problem(wl, dtype=dtype) for wl in wls])
    for wl, r in zip(wls, many):
        try:
            s = solver.fit(W.make_gpu_problem(wl, dtype=dtype))
        except vb.FitError as e:
            s = e.result
        assert r.was_successful() == s.was_successful(), (it, r.minimization_report, s.minimization_report)
        if s.was_successful():
            a, b = np.sort(r.nonlinear_parameters()), np.sort(s.nonlinear_parameters())
            rel = np.max(np.abs(a - b) / np.abs(b))
            worst = max(worst, rel if dtype == np.float64 else 0.0)
            assert rel <= (1e-7 if dtype == np.float64 else 5e-4), (it, dtype, a, b)
        nfits += 1
print(f"stress ok: {rounds} rounds, {nfits} fits, worst fp64 relative parameter difference {worst:.2e}, {time.time()-t0:.1f} s")
