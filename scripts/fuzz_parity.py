"""Randomized parity sweep on the GPU box: random model shapes (tabled and not), sizes, weights and noise levels;
every fit against the CPU oracle (same minimum: residual norm to 1e-9 * ||Y_w||; same success class), fit_many
against fit bitwise. usage: fuzz_parity.py [seed] [cases]"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import workloads as W, varpro_b200 as vb
from test_gpu_round2 import _make_gpu

rng = np.random.default_rng(int(sys.argv[1]) if len(sys.argv) > 1 else 11)
cases = int(sys.argv[2]) if len(sys.argv) > 2 else 40
solver = vb.LevMarSolver.default()
SHAPES = [
    ("exp+1", [(0, [0]), (1, [])], [2.0]),
    ("2exp", [(0, [0]), (0, [1])], [1.0, 5.0]),
    ("2exp+1", [(0, [0]), (0, [1]), (1, [])], [1.0, 4.0]),
    ("3exp", [(0, [0]), (0, [1]), (0, [2])], [0.8, 3.0, 11.0]),
    ("3exp+1", [(0, [0]), (0, [1]), (0, [2]), (1, [])], [0.8, 3.0, 11.0]),
    ("exp+x+1 (off-table)", [(0, [0]), (4, [], 0.5), (1, [])], [2.5]),
    ("4exp+1 (off-table)", [(0, [0]), (0, [1]), (0, [2]), (0, [3]), (1, [])], [0.7, 2.5, 7.0, 20.0]),
]
t0 = time.time()
worst_rn, nfail = 0.0, 0
for c in range(cases):
    name, basis, tau = SHAPES[int(rng.integers(0, len(SHAPES)))]
    m = int(rng.choice([23, 64, 200, 513, 1000, 1024, 1500]))
    S = int(rng.choice([1, 2, 7, 33, 150]))
    x = np.linspace(0.0, 4.0 * max(tau), m)
    cols = []
    for spec in basis:
        if spec[0] == 0: cols.append(np.exp(-x / tau[spec[1][0]]))
        elif spec[0] == 1: cols.append(np.ones_like(x))
        else: cols.append(spec[2] * x)
    Phi = np.stack(cols, axis=1)
    noise = float(rng.choice([1e-2, 1e-4]))
    Y = np.asfortranarray(Phi @ rng.uniform(1.0, 5.0, size=(len(basis), S)) + noise * rng.standard_normal((m, S)))
    w = rng.uniform(0.5, 1.5, size=m) if rng.random() < 0.4 else None
    a0 = list(np.array(tau) * rng.uniform(0.85, 1.2, size=len(tau)))
    wl = dict(x=x, Y=Y, basis=basis, q=len(tau), alpha0=a0, weights=w)
    gp, op = _make_gpu(wl), W.make_oracle(wl)
    try:
        res = solver.fit(gp)
    except vb.FitError as err:  # Err(FitResult) of the reference: unsuccessful termination
        res = err.result
    rep = op.fit()
    Yn = np.linalg.norm(Y if w is None else w[:, None] * Y)
    rn_g, rn_o = np.sqrt(2 * res.minimization_report.objective_function), np.sqrt(2 * rep["objective_function"])
    both_failed = (not res.was_successful()) and (not rep["successful"])
    ok = (res.was_successful() == bool(rep["successful"])) and (both_failed or abs(rn_g - rn_o) <= 1e-9 * Yn)
    worst_rn = max(worst_rn, abs(rn_g - rn_o) / Yn)
    # fit_many of two copies: bitwise the single fit
    g2 = [_make_gpu(wl) for _ in range(2)]
    many = solver.fit_many(g2)
    same = all(np.array_equal(r.nonlinear_parameters(), res.nonlinear_parameters()) and
               r.minimization_report.number_of_evaluations == res.minimization_report.number_of_evaluations for r in many)
    if not ok:
        # same function? evaluate each side's objective at the OTHER side's solution
        a_g, a_o = res.nonlinear_parameters().copy(), np.array(op.params(), dtype=float)
        gp.set_params(a_o)
        rn_g_at_o = np.sqrt(gp.reduce()["rnorm2"])
        op.set_params(a_g)
        r_o = op.residuals()
        rn_o_at_g = np.sqrt(r_o @ r_o) if r_o is not None else float("nan")
        print(f"  objective cross-check: gpu at oracle's alpha {rn_g_at_o:.12e} (oracle {rn_o:.12e}); oracle at gpu's alpha {rn_o_at_g:.12e} (gpu {rn_g:.12e}); alpha_gpu {a_g} alpha_oracle {a_o}")
    if not (ok and same):
        nfail += 1
        print(f"MISMATCH case {c}: {name} m={m} S={S} weighted={w is not None} noise={noise}: gpu ok={res.was_successful()} rn={rn_g:.12e} "
              f"nfev={res.minimization_report.number_of_evaluations} | oracle ok={rep['successful']} rn={rn_o:.12e} | fit_many bitwise={same}")
    for g in g2 + [gp]:
        g.close()
print(f"fuzz: {cases} cases, {nfail} mismatches, worst |rn_gpu - rn_oracle| / ||Y_w|| = {worst_rn:.2e}, {time.time() - t0:.1f} s")
sys.exit(1 if nfail else 0)
