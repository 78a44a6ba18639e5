"""Randomized sweep of the independent-batch kernel: random tabled shapes, sample counts (every row tiling), weights,
batch sizes; every problem against its own vp_fit (same success class, same minimum to 1e-9 * ||y_w||).
usage: fuzz_batch.py [seed] [cases]"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import workloads as W, varpro_b200 as vb
from test_gpu_round2 import _make_gpu
from test_gpu_parity import _batch_model

rng = np.random.default_rng(int(sys.argv[1]) if len(sys.argv) > 1 else 21)
cases = int(sys.argv[2]) if len(sys.argv) > 2 else 30
solver = vb.LevMarSolver.default()
SHAPES = [
    ("exp+1", [(0, [0]), (1, [])], [2.0]),
    ("2exp", [(0, [0]), (0, [1])], [1.0, 5.0]),
    ("2exp+1", [(0, [0]), (0, [1]), (1, [])], [1.0, 4.0]),
    ("3exp", [(0, [0]), (0, [1]), (0, [2])], [0.8, 3.0, 11.0]),
    ("3exp+1", [(0, [0]), (0, [1]), (0, [2]), (1, [])], [0.8, 3.0, 11.0]),
]
t0 = time.time()
nfail, worst = 0, 0.0
for c in range(cases):
    name, basis, tau = SHAPES[int(rng.integers(0, len(SHAPES)))]
    m = int(rng.choice([23, 100, 500, 513, 1000, 1100, 2048, 2500, 4096]))
    if name == "3exp+1" and m > 2048:
        m = 2048  # 7 columns x 4096 rows do not fit in shared memory
    P = int(rng.integers(1, 14))
    x = np.linspace(0.0, 4.0 * max(tau), m)
    Y = np.empty((m, P))
    for p in range(P):
        tp = np.array(tau) * rng.uniform(0.9, 1.1, size=len(tau))
        cols = [np.exp(-x / tp[s[1][0]]) if s[0] == 0 else np.ones_like(x) for s in basis]
        Y[:, p] = np.stack(cols, axis=1) @ rng.uniform(1.0, 5.0, size=len(basis)) + 1e-3 * rng.standard_normal(m)
    Y = np.asfortranarray(Y)
    w = rng.uniform(0.5, 1.5, size=m) if rng.random() < 0.4 else None
    a0 = np.tile(np.array(tau) * rng.uniform(0.9, 1.15, size=len(tau)), (P, 1))
    wl = dict(x=x, basis=basis, q=len(tau))
    batch = vb.IndependentBatch(_batch_model(wl, m), Y, a0, weights=w)
    res = batch.fit()
    for p in range(P):
        one = dict(x=x, Y=np.asfortranarray(Y[:, p:p + 1]), basis=basis, q=len(tau), alpha0=list(a0[p]), weights=w)
        gp = _make_gpu(one)
        try:
            r1 = solver.fit(gp)
        except vb.FitError as err:
            r1 = err.result
        yn = np.linalg.norm(Y[:, p] if w is None else w * Y[:, p])
        d = abs(np.sqrt(2 * res.objective_function[p]) - np.sqrt(2 * r1.minimization_report.objective_function)) / yn
        both_failed = (not res.successful[p]) and (not r1.was_successful())
        if bool(res.successful[p]) != r1.was_successful() or (not both_failed and d > 1e-9):
            nfail += 1
            print(f"MISMATCH case {c} problem {p}: {name} m={m} P={P} weighted={w is not None}: batch ok={bool(res.successful[p])} "
                  f"nfev={res.number_of_evaluations[p]} | fit ok={r1.was_successful()} nfev={r1.minimization_report.number_of_evaluations} | d={d:.2e}")
        elif not both_failed:
            worst = max(worst, d)
        gp.close()
    batch.close()
print(f"fuzz_batch: {cases} cases, {nfail} mismatches, worst relative difference of ||r|| = {worst:.2e}, {time.time() - t0:.1f} s")
sys.exit(1 if nfail else 0)
