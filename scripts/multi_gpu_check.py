"""Multi-GPU check of the column-sharded global fit (run under torchrun, one rank per GPU).

  torchrun --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 scripts/multi_gpu_check.py [S]

Every rank fits its shard of the C2 workload collectively; rank 0 also fits the whole problem on
its own GPU. The sharded result must equal the single-GPU one (parameters to 1e-8 relative; the
sums differ only by summation order), all ranks must agree bitwise, and the gathered coefficients
must match. Prints per-evaluation time of the sharded fit.
"""
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import workloads as W  # noqa: E402
import varpro_b200 as vb  # noqa: E402
from varpro_b200 import sharding  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("gloo")
S = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
wl = W.c2(S=S)
Yl = sharding.shard_observations(wl["Y"], world, rank)
comm = sharding.Communicator(rank, world, device=local)


def build(Y):
    names = ["p0", "p1"]
    model = (vb.SeparableModelBuilder(names).function(["p0"], vb.ExpDecay()).function(["p1"], vb.ExpDecay())
             .invariant_function(vb.Constant()).independent_variable(wl["x"]).initial_parameters(list(wl["alpha0"])).build())
    return vb.SeparableProblemBuilder.mrhs(model).observations(Y).device(local).build()


gp = comm.attach(build(Yl))
red0 = gp.reduce()
dist.barrier()
t0 = time.perf_counter()
res = vb.LevMarSolver.default().fit(gp)
torch.cuda.synchronize()
dt = time.perf_counter() - t0
alpha = res.nonlinear_parameters()
nfev = res.minimization_report.number_of_evaluations
# bitwise agreement of the replicated LM across ranks
al = [None] * world
dist.all_gather_object(al, (alpha.tobytes(), nfev, red0["rnorm2"]))
assert all(a == al[0] for a in al), f"ranks disagree: {al}"
# timing of repeated sharded fits
times = []
for _ in range(5):
    gp.set_params(wl["alpha0"])
    dist.barrier()
    t0 = time.perf_counter()
    r2 = vb.LevMarSolver.default().fit(gp)
    times.append(time.perf_counter() - t0)
C_local = res.linear_coefficients()
Cs = [None] * world
dist.all_gather_object(Cs, C_local)
if rank == 0:
    whole = build(wl["Y"])
    redw = whole.reduce()
    rw = vb.LevMarSolver.default().fit(whole)
    aw = rw.nonlinear_parameters()
    assert abs(red0["rnorm2"] - redw["rnorm2"]) <= 1e-12 * redw["rnorm2"], (red0["rnorm2"], redw["rnorm2"])
    assert np.max(np.abs(red0["H"] - redw["H"])) <= 1e-12 * np.abs(redw["H"]).max()
    assert np.max(np.abs(np.sort(alpha) - np.sort(aw)) / np.abs(np.sort(aw))) <= 1e-8, (alpha, aw)
    assert np.allclose(np.sort(alpha), [1.0, 3.0], atol=1e-8), alpha
    Call = np.concatenate(Cs, axis=1)
    Cw = rw.linear_coefficients()
    if alpha[0] > alpha[1]:
        Call = Call[[1, 0, 2]]
    if aw[0] > aw[1]:
        Cw = Cw[[1, 0, 2]]
    assert np.max(np.abs(Call - wl["C_true"])) <= 1e-6, np.max(np.abs(Call - wl["C_true"]))
    print(f"multi_gpu_check ok: world={world} S={S} alpha={alpha} nfev={nfev} (single GPU nfev={rw.minimization_report.number_of_evaluations}) "
          f"first fit {1e3*dt:.2f} ms, repeat fits {1e3*min(times):.3f} ms = {1e6*min(times)/r2.minimization_report.number_of_evaluations:.1f} us/eval")
dist.barrier()
dist.destroy_process_group()
