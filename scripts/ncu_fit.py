"""One persistent whole-fit launch (fit_kernel_dmma in fit mode) of the canonical C2 problem, for ncu:
launch 0 = problem creation, then (set_params, fit) x 4; the last fit is launch index 8."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import workloads as W
import varpro_b200 as vb
wl = W.c2()
gp = W.make_gpu_problem(wl)
for _ in range(4):
    gp.set_params(wl["alpha0"])
    res = vb.LevMarSolver.default().fit(gp)
print("fit:", res.minimization_report.number_of_evaluations, res.nonlinear_parameters())
