"""K = 1 latency experiment: one C2 fit and one C4 fit, event-timed, for the current context options."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
import workloads as W, varpro_b200 as vb
solver = vb.LevMarSolver.default()
for name, wl, dt in (("C2", W.c2(), np.float64), ("C4", W.c4(), np.float32)):
    gp = W.make_gpu_problem(wl, dtype=dt)
    ext = torch.cuda.ExternalStream(gp._ctx.stream())
    ts = []
    for it in range(12):
        gp.set_params(wl["alpha0"])
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        with torch.cuda.stream(ext):
            e0.record()
        res = solver.fit(gp)
        with torch.cuda.stream(ext):
            e1.record()
        torch.cuda.synchronize()
        if it > 2: ts.append(e0.elapsed_time(e1))
    nf = res.minimization_report.number_of_evaluations
    print(name, "fit_warps", os.environ.get("VP_FIT_WARPS", "8"), "median ms", np.median(ts), "min", np.min(ts), "nfev", nf, "us/eval", 1e3 * np.median(ts) / nf)
    gp.close()
