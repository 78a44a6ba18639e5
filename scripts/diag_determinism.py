"""Diagnostics: is vp_fit_many bitwise vp_fit for every work-item size, and run-to-run reproducible?"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import varpro_b200 as vb  # noqa: E402
import workloads as W  # noqa: E402

sys.path.insert(0, ROOT)
from bench import c2_problem_set  # noqa: E402

K = int(os.environ.get("K", 6))
S = int(os.environ.get("S", 4096))
wls = c2_problem_set(K)
if S != 4096:
    for wl in wls:
        wl["Y"] = np.asfortranarray(wl["Y"][:, :S])
solver = vb.LevMarSolver.default()
seq = [solver.fit(W.make_gpu_problem(wl)) for wl in wls]
ref = [(r.nonlinear_parameters(), r.minimization_report.number_of_evaluations, r.minimization_report.objective_function) for r in seq]
print("vp_fit evaluations:", [r[1] for r in ref])
for ppi in [int(v) for v in os.environ.get("PPI", "1,2,3,5,7,10,37,148,0").split(",")]:
    vb.set_option("queue_parts_per_item", ppi)
    bad = []
    for rep in range(int(os.environ.get("REPS", 4))):
        probs = [W.make_gpu_problem(wl) for wl in wls]
        many = solver.fit_many(probs)
        for k, ((a, nf, ob), b) in enumerate(zip(ref, many)):
            if not (np.array_equal(a, b.nonlinear_parameters()) and nf == b.minimization_report.number_of_evaluations
                    and ob == b.minimization_report.objective_function):
                bad.append((rep, k, nf, b.minimization_report.number_of_evaluations))
        for p in probs:
            p.close()
    print(f"parts_per_item={ppi or 'adaptive'}: {'OK bitwise' if not bad else 'MISMATCH ' + str(bad[:8])}")
