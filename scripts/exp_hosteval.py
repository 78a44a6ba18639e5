"""Host-evaluated (closure) model vs the built-in device model on the C2 shape: per-fit and per-evaluation time."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
import workloads as W, varpro_b200 as vb
from test_gpu_parity import _closure_double_exp_model
wl = W.c2()
solver = vb.LevMarSolver.default()
model = _closure_double_exp_model(wl["x"], wl["alpha0"])
gp_h = vb.SeparableProblemBuilder.mrhs(model).observations(wl["Y"]).build()
gp_d = W.make_gpu_problem(wl)
for name, gp in (("closure model (host-evaluated Phi)", gp_h), ("built-in device model", gp_d)):
    ts = []
    for it in range(6):
        gp.set_params(wl["alpha0"])
        torch.cuda.synchronize(); t0 = time.perf_counter()
        res = solver.fit(gp)
        torch.cuda.synchronize(); ts.append(time.perf_counter() - t0)
    nf = res.minimization_report.number_of_evaluations
    print(f"{name}: {1e3*np.median(ts[1:]):.3f} ms per fit, {nf} evaluations, {1e6*np.median(ts[1:])/nf:.1f} us per evaluation, alpha {res.nonlinear_parameters()}")
