"""Diagnostics: BASELINE config 3 problems whose GPU fit was unsuccessful -- what does the oracle do from the same start?"""
import os, sys
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import varpro_b200 as vb
import workloads as W
P, m = int(os.environ.get("P", 65536)), 4096
rng = np.random.Generator(np.random.PCG64(65536))
tau = np.array([1.0, 3.0, 9.0]) * rng.uniform(0.8, 1.25, size=(P, 3))
c = rng.uniform(1.0, 10.0, size=(P, 3))
gen = torch.Generator(device="cuda"); gen.manual_seed(65536)
x = np.linspace(0.0, 20.0, m); xd = torch.from_numpy(x).cuda()
Yd = torch.empty((P, m), dtype=torch.float64, device="cuda")
for b0 in range(0, P, 8192):
    t = torch.from_numpy(tau[b0:b0 + 8192]).cuda(); cc = torch.from_numpy(c[b0:b0 + 8192]).cuda()
    blk = sum(cc[:, j:j + 1] * torch.exp(-xd[None, :] / t[:, j:j + 1]) for j in range(3))
    blk += 1e-3 * torch.randn(blk.shape, generator=gen, device="cuda", dtype=torch.float64)
    Yd[b0:b0 + 8192] = blk
alpha0 = tau * np.array([1.3, 0.8, 1.2])
names = ["p0", "p1", "p2"]
b = vb.SeparableModelBuilder(names)
for k in range(3):
    b = b.function([names[k]], vb.ExpDecay())
model = b.independent_variable(x).initial_parameters([1.0, 1.0, 1.0]).build()
batch = vb.IndependentBatch(model, None, alpha0, y_device_ptr=Yd.data_ptr(), P=P)
res = batch.fit()
ok = res.successful
from collections import Counter
print("GPU terminations:", Counter(str(t) for t in res.terminations))
bad = np.flatnonzero(~ok)
print("unsuccessful:", len(bad), "evals of those: mean", res.number_of_evaluations[bad].mean(), "max", res.number_of_evaluations[bad].max())
Yh = Yd[torch.as_tensor(bad[:40], device="cuda")].cpu().numpy()
for row, p in enumerate(bad[:40]):
    one = dict(x=x, Y=np.asfortranarray(Yh[row][:, None]), basis=W.TRIPLE_EXP, q=3, alpha0=list(alpha0[p]), weights=None)
    op = W.make_oracle(one); rep = op.fit()
    print(p, "GPU", res.terminations[p], res.number_of_evaluations[p], "%.6e" % np.sqrt(2*res.objective_function[p]), np.sort(res.nonlinear_parameters[p]),
          "| oracle", rep["termination"], rep["number_of_evaluations"], "%.6e" % np.sqrt(2*rep["objective_function"]), np.sort(op.params()))

print("---- class-b problems through the single-problem path and step by step")
sys.path.insert(0, os.path.join(ROOT, "tests"))
import lm_harness as LH
solver = vb.LevMarSolver.default()
shown = 0
for row, p in enumerate(bad[:40]):
    one = dict(x=x, Y=np.asfortranarray(Yh[row][:, None]), basis=W.TRIPLE_EXP, q=3, alpha0=list(alpha0[p]), weights=None)
    op = W.make_oracle(one); rep = op.fit()
    if not rep["successful"]:
        continue
    gp = W.make_gpu_problem(one)
    try:
        r1 = solver.fit(gp)
        t1 = (str(r1.minimization_report.termination), r1.minimization_report.number_of_evaluations)
    except vb.FitError as e:
        t1 = ("FitError " + str(e.result.minimization_report.termination), e.result.minimization_report.number_of_evaluations)
    print(p, "single-problem GPU path:", t1, "| oracle", rep["termination"], rep["number_of_evaluations"])
    # step by step: product LM state machine fed with GPU evaluations; oracle evaluated at the same points
    gp2 = W.make_gpu_problem(one); op2 = W.make_oracle(one)
    h = LH.LmHarness(alpha0[p]); xx = np.asarray(alpha0[p], dtype=np.float64); more = True; it = 0
    while more and it < 6:
        gp2.set_params(xx); red = gp2.reduce(); ok_o = op2.set_params(xx); r_o = op2.residuals()
        print("   eval", it, "x", xx, "GPU finite", None if red is None else red["finite"], "rnorm2", None if red is None else "%.4e" % red["rnorm2"],
              "| oracle ok", ok_o, "rnorm2", None if r_o is None else "%.4e" % (r_o @ r_o))
        if red is None:
            break
        more = h.advance(red["rnorm2"], red["g"], red["H"], finite=red["finite"]); xx = h.trial(); it += 1
    shown += 1
    if shown >= 4:
        break
