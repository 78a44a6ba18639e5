#!/usr/bin/env python
"""bench.py -- fits/sec of the VarPro hot path on BASELINE.json's headline workload.

Workload (config.workload = "C2"): double-exponential + offset, 1024 samples, MRHS S = 4096
right-hand sides with shared nonlinear parameters (global fit), fp64 -- BASELINE.json configs[1],
the B200 restatement of the reference's benches/multiple_right_hand_sides.rs (which uses S=1000).
One step = one complete LevMarSolver::fit from the stated initial guess (2, 6.5) to LM
convergence with the crate-default tolerances, starting from a built problem (the reference's
criterion bench builds the problem in its un-timed setup closure and times `fit`).

  value : whole-job fits/s with the observations already resident in HBM when the timed region
          starts. K distinct problem instances (K x 33.5 MB, more than the 126 MB L2 for K >= 4)
          are built un-timed; the timed region is exactly K fits made by ONE vp_fit_many call (all
          K fits share one persistent grid through a device-side work queue), CUDA events on the
          library's stream, barrier + synchronize on both sides, max over ranks. `latency_mode`
          reports the same K fits made one after the other (vp_fit, whole GPU per fit).
  e2e   : the same metric through the reference-facing API with HOST buffers: every step builds
          the problem from pinned host memory (H2D of Y inside the timed region), fits, and
          reads parameters + linear coefficients back (D2H); a few host threads pipeline steps so
          that one step's copy overlaps another step's fit (PCIe-bound: 33.6 MB per step).
  roofline : the dominant kernel of the timed region (fit_queue_kernel): algorithmic bytes = 8*m*S per
          evaluation x the evaluations made in the timed region, over the region's CUDA-event time.
          `single_evaluation_full_grid` is one fused evaluation launch on the whole GPU with L2
          flushed before every launch (and un-flushed).
  cpu_baseline : the CPU restatement of the reference algorithm (oracle/, "port"), 1 thread
          (the reference is single-threaded), on a bounded sample of the same workload.

`--impl reference` times that CPU restatement with all host threads instead (the Rust crate
cannot be built in this image: no cargo/rustc).

N > 1 (torchrun): every rank fits its own copy of the same K independent problems (weak scaling: fixed
per-GPU work, no data-path collective); value = total fits / max-over-ranks time. A `sharded_global_fit`
object reports BASELINE config 5 (one global fit, 131 072 columns per GPU, in-kernel NVLink exchange).
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

M, S_C2, N_BASIS, Q = 1024, 4096, 3, 2
BYTES_EVAL = 8 * M * S_C2  # algorithmic bytes of one evaluation: each weighted observation read once


def c2_workload(S=S_C2, seed=2314093240213841123 % (2 ** 63)):
    import workloads as W
    return W.c2(S=S, seed=seed)


# ---------------------------------------------------------------------------------------------
# clocks sampling (B200_PROFILING.md recipe)
# ---------------------------------------------------------------------------------------------
class ClockSampler:
    """One long-running `nvidia-smi -lms` child started before the timed regions and stopped after them
    (the recipe in B200_PROFILING.md): no per-sample fork from this process, whose host thread is
    inside the timed region."""
    FIELDS = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
              "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index, enabled=True, period_ms=25):
        self.index = index
        self.enabled = enabled
        self.period_ms = period_ms
        self.samples = []
        self.reasons = set()
        self._p = None

    def __enter__(self):
        if self.enabled:
            try:
                self._p = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.FIELDS}",
                                            "--format=csv,noheader,nounits", "-lms", str(self.period_ms)],
                                           stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            except Exception:
                self._p = None
        return self

    def __exit__(self, *a):
        if self._p is None:
            return
        try:
            self._p.terminate()
            out, _ = self._p.communicate(timeout=5)
        except Exception:
            out = ""
        for line in out.splitlines():
            p = [x.strip() for x in line.split(",")]
            if len(p) < 6:
                continue
            try:
                self.samples.append((float(p[0]), float(p[1])))
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), p[2:6]):
                if val.lower().startswith("active"):
                    self.reasons.add(name)
        self._p = None

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": sorted(self.reasons)}
        sm = sorted(s[0] for s in self.samples)
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": self.samples[0][1], "reasons": sorted(self.reasons),
                "samples": len(sm)}


# ---------------------------------------------------------------------------------------------
# CPU arm: the oracle port (test infrastructure; the one place bench.py executes oracle/)
# ---------------------------------------------------------------------------------------------
def cpu_fit_seconds(wl, threads):
    import workloads as W
    from oracle import varpro_oracle as vo
    vo.set_threads(threads)
    op = W.make_oracle(wl)  # build is un-timed (criterion setup closure)
    t0 = time.perf_counter()
    rep = op.fit()
    dt = time.perf_counter() - t0
    assert rep["successful"], rep
    return dt, rep


def run_reference(args, rank, world):
    """--impl reference: CPU restatement of the reference algorithm, all host threads, rank 0 only."""
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    wl = c2_workload()
    for _ in range(min(args.warmup, 1)):
        cpu_fit_seconds(wl, cores)
    steps = max(1, min(args.steps, 5))  # bounded: one C2 fit is ~2 s of CPU work
    t = [cpu_fit_seconds(wl, cores)[0] for _ in range(steps)]
    v = steps / sum(t)
    line = {
        "impl": "reference", "metric": "fits/sec (double-exp MRHS, 1024 samples)", "value": v, "unit": "fits/s",
        "n_gpus": args.gpus, "steps": steps, "warmup": min(args.warmup, 1), "ms_per_step": 1e3 * sum(t) / steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": "C2", "m": M, "S": S_C2, "n": N_BASIS, "q": Q, "alpha0": [2.0, 6.5]},
        "cpu_baseline": {"value": v, "unit": "fits/s", "cores": cores, "kind": "port",
                         "sample": f"{steps} complete C2 fits (S={S_C2}) with the C restatement of varpro 0.13.3 "
                                   "(oracle/varpro_oracle.c, OpenMP over right-hand sides); not the Rust binary"},
        "e2e": {"value": v, "unit": "fits/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


# ---------------------------------------------------------------------------------------------
# GPU arm
# ---------------------------------------------------------------------------------------------
def run_gpu(args, rank, world, local_rank):
    import torch
    import torch.distributed as dist

    import varpro_b200 as vb
    import workloads as W
    from varpro_b200 import _lib, api

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (the product path has no CPU fallback)")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    lib = _lib.load()
    K, Wm = args.steps, max(args.warmup, 3)
    wl = c2_workload()  # every rank fits the same set of problems: fixed per-GPU work (weak scaling)
    solver = vb.LevMarSolver.default()

    def build(device_problem=True):
        import workloads as W2
        return _build_on_device(W2, wl, local_rank)

    ctx = api._Ctx.get(local_rank)
    ext = torch.cuda.ExternalStream(ctx.stream(), device=local_rank)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- value: observations resident in HBM, K distinct built problems, fitted concurrently ----
    # (throughput mode, vp_fit_many: one persistent grid serves all K fits through a device-side work queue)
    def timed_fits(probs, many):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        with torch.cuda.stream(ext):
            e0.record()
        if many:
            res = solver.fit_many(probs)
        else:
            res = [solver.fit(p) for p in probs]
        with torch.cuda.stream(ext):
            e1.record()
        barrier()
        assert all(r.was_successful() for r in res)
        return e0.elapsed_time(e1), [r.minimization_report.number_of_evaluations for r in res]

    def fresh(n):
        return [build() for _ in range(n)]

    for _ in range(Wm):  # warm-up steps: whole batches (also builds the side streams, loads the kernels)
        wp = fresh(K)
        solver.fit_many(wp)
        for p in wp:
            p.close()
    probs = fresh(K)
    launches0 = ctx.kernel_launches()
    clocks = ClockSampler(local_rank, enabled=(rank == 0))  # one sampler per job: nvidia-smi polling costs host CPU
    clocks.__enter__()  # samples until the end of the e2e region: all three timed regions run under it
    ms, nfev = timed_fits(probs, many=True)
    launches = ctx.kernel_launches() - launches0
    alpha = np.sort(probs[-1].params())
    assert np.allclose(alpha, [1.0, 3.0], atol=1e-8), alpha
    t = torch.tensor([ms], dtype=torch.float64, device="cuda")
    rank_ms = [ms]
    if world > 1:
        allt = [torch.zeros_like(t) for _ in range(world)]
        dist.all_gather(allt, t)
        rank_ms = [float(v.item()) for v in allt]
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_max = float(t.item())
    timed_evals = int(sum(nfev) - len(nfev))  # the evaluation at the starting point belongs to the (un-timed) build
    for p in probs:
        p.close()
    # latency mode for comparison: the same K fits one after the other, each on the whole GPU
    probs = fresh(K)
    ms_seq, nfev_seq = timed_fits(probs, many=False)
    for p in probs:
        p.close()
    probs.clear()

    # ---- e2e: host buffers -> build -> fit -> read back, every step ------------------------
    # A few host threads (one library context = one stream each) each take chunks of steps: build the
    # chunk's problems from pinned host memory (one H2D of Y per step), fit them together
    # (vp_fit_many), read every step's parameters and coefficients back (D2H). The copies of one
    # thread overlap the fits of another; every step still pays its own copies.
    NTH = int(os.environ.get("VP_E2E_THREADS", "3"))
    CH = int(os.environ.get("VP_E2E_CHUNK", "1"))  # steps a worker builds, fits together (vp_fit_many) and reads back
    Yh = [torch.from_numpy(np.ascontiguousarray(wl["Y"].T)).pin_memory() for _ in range(NTH)]  # (S, m) row-major == m x S col-major
    Yv = [y.numpy().T for y in Yh]  # Fortran-ordered views of the pinned buffers
    h2d = Yh[0].numel() * 8 + M * 8
    d2h = N_BASIS * S_C2 * 8 + Q * 8

    def e2e_chunk(slot, nsteps):
        ps = [W.make_gpu_problem(wl, Y=Yv[slot], device=local_rank, ctx_slot=1 + slot) for _ in range(nsteps)]  # H2D per step
        rs = solver.fit_many(ps)
        out = [(r.nonlinear_parameters(), r.linear_coefficients()) for r in rs]  # D2H per step
        for p in ps:
            p.close()
        return out

    from concurrent.futures import ThreadPoolExecutor
    pool = ThreadPoolExecutor(NTH)

    def e2e_run(nsteps):
        chunks = [min(CH, nsteps - i) for i in range(0, nsteps, CH)]
        futs = [pool.submit(e2e_chunk, i % NTH, c) for i, c in enumerate(chunks)]
        return [o for f in futs for o in f.result()]

    e2e_run(max(Wm, NTH * CH))
    barrier()
    t0 = time.perf_counter()
    outs = e2e_run(K)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    assert len(outs) == K
    a, c = outs[-1]
    clocks.__exit__(None, None, None)
    pool.shutdown()
    t = torch.tensor([dt], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_val = world * K / float(t.item())
    assert np.allclose(np.sort(a), [1.0, 3.0], atol=1e-8)

    # ---- N > 1: ONE global fit column-sharded over the ranks (BASELINE config 5 shape) -----------
    sharded = None
    if world > 1:
        sharded = _sharded_global_fit(torch, dist, vb, api, solver, wl, rank, world, local_rank)

    # ---- roofline of the streaming kernel (rank 0) -------------------------------------------
    line = None
    if rank == 0:
        p = build()
        pu, su = C.c_double(), C.c_double()
        g, sm = C.c_int64(), C.c_int64()
        lib.vp_profile_evaluation(p._h, 50, 0, C.byref(pu), C.byref(su), C.byref(g), C.byref(sm))
        warm_us, panel_us = su.value, pu.value
        lib.vp_profile_evaluation(p._h, 50, 512 << 20, C.byref(pu), C.byref(su), C.byref(g), C.byref(sm))
        cold_us = su.value
        p.close()
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        peak = float(peaks.get("hbm_gbs", 6650.0))
        # dominant kernel = the persistent fit kernel; K launches run concurrently in the timed region
        achieved = timed_evals * BYTES_EVAL / (ms * 1e-3) / 1e9
        single_cold = BYTES_EVAL / (cold_us * 1e-6) / 1e9
        traffic = None
        try:
            traffic = json.load(open(os.path.join(ROOT, "profiles", "r01_queue_traffic.json")))["dram_bytes_per_launch"]
        except Exception:
            pass
        # CPU baseline: oracle port, 1 thread, bounded sample (rank 0, N = 1 only)
        cpu = None
        if world == 1 and not args.no_cpu:
            wl_cpu = c2_workload()
            NCPU = 10  # bounded sample: ~10 s of single-thread CPU work
            runs = [cpu_fit_seconds(wl_cpu, 1) for _ in range(NCPU)]
            dtc = sum(r[0] for r in runs)
            rep = runs[-1][1]
            cpu = {"value": NCPU / dtc, "unit": "fits/s", "cores": 1, "kind": "port",
                   "sample": f"{NCPU} complete C2 fits (S={S_C2}, {rep['number_of_evaluations']} residual + "
                             f"{rep['number_of_jacobians']} Jacobian evaluations each, {dtc:.1f} s in total) with the C "
                             "restatement of varpro 0.13.3 (oracle/varpro_oracle.c), single thread like the reference; "
                             "not the Rust binary"}
        fits = world * K
        line = {
            "metric": "fits/sec (double-exp MRHS, 1024 samples)", "value": fits / (ms_max * 1e-3), "unit": "fits/s",
            "n_gpus": world, "steps": K, "warmup": Wm, "ms_per_step": ms_max / K, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": "C2", "m": M, "S": S_C2, "n": N_BASIS, "q": Q, "alpha0": [2.0, 6.5],
                       "problems_per_rank": K, "concurrent_fits": K, "rank_ms": rank_ms,
                       "l2_policy": "K distinct 33.5 MB problems are fitted concurrently: K*33.5 MB of inputs are streamed "
                                    "per evaluation round (larger than the 126 MB L2 for K >= 4)",
                       "evals_per_fit_mean": float(np.mean(nfev)), "fit_mode": os.environ.get("VP_FIT_MODE", "persistent")},
            "clocks": clocks.summary(),
            "e2e": {"value": e2e_val, "unit": "fits/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "host_threads": NTH, "steps_per_chunk": CH},
            "gpu_launches": int(launches),
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": traffic, "kernel": "fit_queue_kernel<3,2,32,8> (all K fits on one persistent grid: work queue of "
                                                      "(fit, chunk) items; panel + Y-streaming reduce + LM step in-kernel)",
                         "how": "ONE launch fits all K problems (vp_fit_many): achieved = algorithmic bytes of the launch "
                                "(timed evaluations x 8*m*S) / CUDA-event time of the timed region",
                         "bytes_per_launch": BYTES_EVAL * timed_evals, "timed_evaluations": timed_evals,
                         "fits_per_launch": K, "region_us": ms * 1e3,
                         "single_evaluation_full_grid": {"us_l2_flushed": cold_us, "GBps_l2_flushed": single_cold,
                                                         "frac_l2_flushed": single_cold / peak, "us_l2_warm": warm_us,
                                                         "GBps_l2_warm": BYTES_EVAL / (warm_us * 1e-6) / 1e9,
                                                         "grid": int(g.value), "smem_bytes": int(sm.value)},
                         "peak_source": "MEASURED_PEAKS.json hbm_gbs" if peaks else "fallback 6650 GB/s"},
            "latency_mode": {"value": world * K / (ms_seq * 1e-3), "unit": "fits/s", "ms_per_fit": ms_seq / K,
                             "what": "the same K fits one after the other (vp_fit), each on the whole GPU",
                             "evals_per_fit_mean": float(np.mean(nfev_seq))},
            "cpu_baseline": cpu,
        }
        if sharded is not None:
            line["sharded_global_fit"] = sharded
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def _sharded_global_fit(torch, dist, vb, api, solver, wl, rank, world, device, cols_per_rank=131072, reps=5):
    """BASELINE config 5: double-exponential MRHS with 131 072 columns per GPU (1.07 GB fp64 each), ONE global
    fit whose (||r||^2, J^T r, J^T J) are exchanged through the NVLink mailboxes inside the fit kernel."""
    from varpro_b200 import sharding
    x = torch.from_numpy(np.asarray(wl["x"], dtype=np.float64)).cuda()
    gen = torch.Generator(device="cuda")
    gen.manual_seed(1048576 + rank)
    Cs = torch.rand(3, cols_per_rank, generator=gen, device="cuda", dtype=torch.float64) * 100.0
    Phi = torch.stack([torch.exp(-x / 1.0), torch.exp(-x / 3.0), torch.ones_like(x)], dim=1)
    Yd = (Phi @ Cs).T.contiguous()  # (S_local, m) row-major == m x S_local column-major
    torch.cuda.synchronize()
    names = ["p0", "p1"]
    model = (vb.SeparableModelBuilder(names).function(["p0"], vb.ExpDecay()).function(["p1"], vb.ExpDecay())
             .invariant_function(vb.Constant()).independent_variable(wl["x"]).initial_parameters(list(wl["alpha0"])).build())
    p = api.SeparableProblem(model, None, None, -1.0, False, device, y_device_ptr=Yd.data_ptr(), S=cols_per_rank, ldY=M)
    del Yd
    comm = sharding.Communicator(rank, world, device=device)
    comm.attach(p)
    times, nfev = [], 0
    for it in range(reps + 1):
        p.set_params(wl["alpha0"])  # collective evaluation (un-timed): back to the starting point
        torch.cuda.synchronize()
        dist.barrier()
        t0 = time.perf_counter()
        r = solver.fit(p)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        nfev = r.minimization_report.number_of_evaluations
        if it > 0:
            times.append(dt)
    alpha = np.sort(r.nonlinear_parameters())
    assert np.allclose(alpha, [1.0, 3.0], atol=1e-8), alpha
    t = torch.tensor([min(times)], dtype=torch.float64, device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    best = float(t.item())
    p.close()
    comm.close()
    S_total = cols_per_rank * world
    return {"workload": "C5 shape: double-exp MRHS, one global fit, columns sharded over the ranks",
            "S_total": S_total, "columns_per_gpu": cols_per_rank, "fits_per_s": 1.0 / best, "ms_per_fit": 1e3 * best,
            "evaluations": nfev, "us_per_evaluation": 1e6 * best / max(nfev - 1, 1),
            "aggregate_GBps": 8.0 * M * S_total * max(nfev - 1, 1) / best / 1e9,
            "collective": "in-kernel NVLink mailbox exchange of (||r||^2, J^T r, J^T J), one per evaluation; no NCCL on the data path"}


def _build_on_device(W, wl, device):
    """Build a problem whose observations are already on the device (vp_problem_create_device)."""
    import torch

    import varpro_b200 as vb
    from varpro_b200 import api
    names = ["p0", "p1"]
    model = (vb.SeparableModelBuilder(names).function(["p0"], vb.ExpDecay()).function(["p1"], vb.ExpDecay())
             .invariant_function(vb.Constant()).independent_variable(wl["x"]).initial_parameters(list(wl["alpha0"])).build())
    Yd = torch.from_numpy(np.ascontiguousarray(wl["Y"].T)).to(f"cuda:{device}")  # (S, m) row-major == m x S col-major
    torch.cuda.synchronize()
    p = api.SeparableProblem(model, None, None, -1.0, False, device, y_device_ptr=Yd.data_ptr(), S=wl["Y"].shape[1],
                             ldY=wl["Y"].shape[0])
    del Yd
    return p


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return
    run_gpu(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
