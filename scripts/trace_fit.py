import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np, workloads as W, varpro_b200 as vb
wl = W.c2(S=int(os.environ.get("S", 4096)))
gp = W.make_gpu_problem(wl)
res = vb.LevMarSolver.default().fit(gp)
print(res.minimization_report, res.nonlinear_parameters())
