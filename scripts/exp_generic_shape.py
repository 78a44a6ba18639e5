"""How fast is the generic path (model shapes without a compiled fast path)? Four exponentials + offset (n = 5, q = 4),
m = 1024, S = 4096, against the double-exponential + offset of the same size on the fused kernels."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
import workloads as W, varpro_b200 as vb
from test_gpu_round2 import _make_gpu, FOUR_EXP
rng = np.random.default_rng(3)
m, S = 1024, 4096
x = np.linspace(0.0, 30.0, m)
tau = np.array([0.7, 2.5, 7.0, 20.0])
Phi = np.stack([np.exp(-x / t) for t in tau] + [np.ones_like(x)], axis=1)
Y = np.asfortranarray(Phi @ rng.uniform(1.0, 5.0, size=(5, S)) + 1e-4 * rng.standard_normal((m, S)))
wl = dict(x=x, Y=Y, basis=FOUR_EXP, q=4, alpha0=list(tau * np.array([1.15, 0.9, 1.1, 0.92])), weights=None)
solver = vb.LevMarSolver.default()
for name, w_ in (("4 exp + offset (generic path)", wl), ("2 exp + offset (fused path)", W.c2())):
    gp = _make_gpu(w_) if name.startswith("4") else W.make_gpu_problem(w_)
    ts = []
    for it in range(6):
        gp.set_params(w_["alpha0"])
        torch.cuda.synchronize(); t0 = time.perf_counter()
        res = solver.fit(gp)
        torch.cuda.synchronize(); ts.append(time.perf_counter() - t0)
    nf = res.minimization_report.number_of_evaluations
    print(f"{name}: {1e3*np.median(ts[1:]):.3f} ms per fit, {nf} evaluations, {1e6*np.median(ts[1:])/nf:.1f} us per evaluation, success {res.was_successful()}")
    gp.close()
