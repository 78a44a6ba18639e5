#!/usr/bin/env python
"""bench.py -- fits/sec of the VarPro hot path on BASELINE.json's headline workload.

Workload (config.workload = "C2"): double-exponential + offset, 1024 samples, MRHS S = 4096
right-hand sides with shared nonlinear parameters (global fit), fp64 -- BASELINE.json configs[1],
the B200 restatement of the reference's benches/multiple_right_hand_sides.rs (which uses S=1000).
One step = one complete LevMarSolver::fit from the stated initial guess (2, 6.5) to LM
convergence with the crate-default tolerances, starting from a built problem (the reference's
criterion bench builds the problem in its un-timed setup closure and times `fit`).

  value : whole-job fits/s with the observations already resident in HBM when the timed region
          starts. K = --steps DISTINCT problems (problem 0 is the canonical C2 instance; problem k > 0 has its
          own seed, its own true parameters tau* = (1, 3) * U[0.85, 1.15) and its own coefficients, so the fits
          walk different LM paths and need different numbers of evaluations) are built un-timed; the timed
          region is exactly K fits made by ONE vp_fit_many call (one persistent grid, device-side work queue),
          CUDA events on the library's stream, barrier + synchronize on both sides, max over ranks. The
          launch is repeated (config.timed_launches, the problems reset to the initial guess un-timed in
          between); value comes from the MEDIAN launch, min / max are reported next to it.
          `by_concurrency` repeats this for K = 1 (one fit on the whole GPU: vp_fit, L2 flushed before
          every launch), K = --steps and K = 60; `latency_mode` is the K fits made one after the other.
  e2e   : the same metric through the reference-facing API with HOST buffers (LevMarSolver.fit_host_batch =
          vp_fit_host_batch): every step builds the problem from pinned host memory (H2D of Y inside the
          timed region), fits, and reads parameters + linear coefficients back (D2H); three worker threads
          of the library pipeline steps so that one step's copy overlaps another step's fit (PCIe-bound:
          33.6 MB per step). The line
          carries the achieved H2D GB/s per GPU next to the raw pinned-memcpy bandwidth measured in
          the same run with all ranks copying at once (the ceiling of this number).
  roofline : the dominant kernel of the timed region (fit_queue_kernel): algorithmic bytes = 8*m*S per
          evaluation x the evaluations made in the timed launch, over that launch's CUDA-event time.
  extra : BASELINE configs 3 and 4 at full size, each with its own roofline object (N = 1, rank 0).
  cpu_baseline : the CPU restatement of the reference algorithm (oracle/, "port"), 1 thread
          (the reference is single-threaded), on a bounded sample of the same workload.

`--impl reference` times that CPU restatement with all host threads instead (the Rust crate
cannot be built in this image: no cargo/rustc).

N > 1 (torchrun): every rank fits its own K problems (weak scaling: fixed per-GPU work, no data-path
collective); value = total fits / max-over-ranks time. Two more objects report BASELINE config 5 in both of its
readings: `sharded_global_fit` (ONE global fit, 131 072 columns per GPU, in-kernel NVLink exchange; checked
against the generating coefficients) and `independent_batch_c5` (131 072 independent double-exponential problems
per GPU through vp_batch_*, no collective).
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

M, S_C2, N_BASIS, Q = 1024, 4096, 3, 2
BYTES_EVAL = 8 * M * S_C2  # algorithmic bytes of one evaluation: each weighted observation read once
C2_SEED = 2314093240213841123 % (2 ** 63)
METRIC = "fits/sec (double-exp MRHS, 1024 samples)"


def c2_workload(S=S_C2, seed=C2_SEED):
    import workloads as W
    return W.c2(S=S, seed=seed)


def c2_problem_set(K):
    """K distinct C2-shaped problems. k = 0: the canonical instance (tau* = (1, 3), the bench's seed).
    k > 0: own seed, tau* = (1, 3) * U[0.85, 1.15), C* ~ U[0, 100): noise-free like the reference bench."""
    import workloads as W
    out = []
    for k in range(K):
        wl = W.c2(S=S_C2, seed=C2_SEED + k)
        if k > 0:
            rng = np.random.Generator(np.random.PCG64(900000 + k))
            tau = np.array([1.0, 3.0]) * rng.uniform(0.85, 1.15, size=2)
            x = wl["x"]
            Phi = np.stack([np.exp(-x / tau[0]), np.exp(-x / tau[1]), np.ones_like(x)], axis=1)
            wl["Y"] = np.asfortranarray(Phi @ wl["C_true"])
            wl["alpha_true"] = list(tau)
        out.append(wl)
    return out


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
        "impl": "reference", "metric": METRIC, "value": v, "unit": "fits/s",
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
def _stats(xs):
    xs = sorted(xs)
    return {"median": xs[len(xs) // 2], "min": xs[0], "max": xs[-1], "n": len(xs)}


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
    R = max(args.repeats, 1)
    solver = vb.LevMarSolver.default()
    ctx = api._Ctx.get(local_rank)
    ext = torch.cuda.ExternalStream(ctx.stream(), device=local_rank)
    flush_buf = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")  # > 126 MB L2

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(v):
        t = torch.tensor([v], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---- the problem set: K_max distinct problems, built once (un-timed), reset between launches ----
    K_big = 60 if (world == 1 and not args.quick) else K
    K_max = max(K, K_big)
    wls = c2_problem_set(K_max)
    probs = [_build_on_device(wl, local_rank) for wl in wls]
    alpha0 = list(wls[0]["alpha0"])

    def reset(ps):
        for p in ps:
            p.set_params(alpha0)  # one evaluation at the starting point: the built state of the reference's bench

    def timed(ps, many, flush=False):
        """One timed launch: (ms, evaluations per fit)."""
        reset(ps)
        if flush:
            flush_buf.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        with torch.cuda.stream(ext):
            e0.record()
        if many:
            res = solver.fit_many(ps)
        else:
            res = [solver.fit(p) for p in ps]
        with torch.cuda.stream(ext):
            e1.record()
        barrier()
        assert all(r.was_successful() for r in res)
        return e0.elapsed_time(e1), [r.minimization_report.number_of_evaluations for r in res]

    def check_truth(ps, wls_):
        for p, wl in zip(ps, wls_):
            a = np.sort(p.params())
            t = np.sort(np.asarray(wl["alpha_true"], dtype=np.float64))
            assert np.max(np.abs(a - t) / t) <= 1e-8, (a, t)

    for _ in range(Wm):  # warm-up steps: whole batches (loads the kernels, ramps the clocks)
        reset(probs[:K])
        solver.fit_many(probs[:K])
    clocks = ClockSampler(local_rank, enabled=(rank == 0))  # one sampler per job: nvidia-smi polling costs host CPU
    clocks.__enter__()  # samples until the end of the e2e region: all timed regions run under it
    launches0 = ctx.kernel_launches()
    runs = [timed(probs[:K], many=True) for _ in range(R)]
    launches = (ctx.kernel_launches() - launches0 - R * K) // R  # minus the un-timed resets (one launch each)
    check_truth(probs[:K], wls[:K])
    ms_list = [r[0] for r in runs]
    nfev = runs[0][1]
    assert all(r[1] == nfev for r in runs), "evaluation counts must not depend on the run (deterministic partial sums)"
    ms_med_local = _stats(ms_list)["median"]
    rank_ms = [ms_med_local]
    if world > 1:
        t = torch.tensor([ms_med_local], dtype=torch.float64, device="cuda")
        allt = [torch.zeros_like(t) for _ in range(world)]
        dist.all_gather(allt, t)
        rank_ms = [float(v.item()) for v in allt]
    ms_med = max_over_ranks(ms_med_local)
    ms_min, ms_max = max_over_ranks(min(ms_list)), max_over_ranks(max(ms_list))
    timed_evals = int(sum(nfev) - len(nfev))  # the evaluation at the starting point belongs to the (un-timed) build

    # ---- the same metric at other concurrencies: K = 1 (one fit on the whole GPU), K = 60 ----------
    by_k = {}
    one = [timed(probs[:1], many=False, flush=True) for _ in range(R)]
    by_k["1"] = {"fits_per_s": world * 1e3 / max_over_ranks(_stats([r[0] for r in one])["median"]),
                 "ms": _stats([r[0] for r in one]), "evaluations": one[0][1][0],
                 "roofline_frac": None, "l2": "flushed before every launch (the fit itself re-reads its 33.5 MB from L2)"}
    by_k[str(K)] = {"fits_per_s": world * K * 1e3 / ms_med, "ms": _stats(ms_list), "evaluations_mean": float(np.mean(nfev))}
    if K_big > K:
        big = [timed(probs[:K_big], many=True) for _ in range(max(3, R // 2))]
        check_truth(probs[:K_big], wls[:K_big])
        by_k[str(K_big)] = {"fits_per_s": world * K_big * 1e3 / max_over_ranks(_stats([r[0] for r in big])["median"]),
                            "ms": _stats([r[0] for r in big]), "evaluations_mean": float(np.mean(big[0][1])),
                            "timed_evaluations": int(sum(big[0][1]) - K_big)}
    # latency mode for comparison: the same K fits one after the other, each on the whole GPU
    seq = [timed(probs[:K], many=False) for _ in range(3)]
    ms_seq = max_over_ranks(_stats([r[0] for r in seq])["median"])
    nfev_seq = seq[0][1]
    assert nfev_seq == nfev, "vp_fit and vp_fit_many must need the same number of evaluations (bitwise-equal partial sums)"
    for p in probs[K:]:
        p.close()
    del probs[K:]

    # ---- e2e: host buffers -> build -> fit -> read back, every step ------------------------
    # A few host threads (one library context = one stream each) each take steps: build the problem from pinned
    # host memory (one H2D of Y per step), fit, read the step's parameters and coefficients back (D2H). The
    # copies of one thread overlap the fits of another; every step still pays its own copies.
    NTH = int(os.environ.get("VP_E2E_THREADS", "3"))
    Yh = [torch.from_numpy(np.ascontiguousarray(wls[i % K]["Y"].T)).pin_memory() for i in range(NTH)]  # (S, m) row-major == m x S col-major
    Yv = [y.numpy().T for y in Yh]  # Fortran-ordered views of the pinned buffers
    h2d = Yh[0].numel() * 8 + M * 8
    d2h = N_BASIS * S_C2 * 8 + Q * 8
    # the ceiling: raw pinned host-to-device copies of the same size, all ranks at once
    dst = torch.empty_like(Yh[0], device="cuda")
    barrier()
    t0 = time.perf_counter()
    for _ in range(8):
        dst.copy_(Yh[0], non_blocking=True)
    torch.cuda.synchronize()
    raw_gbps = 8 * Yh[0].numel() * 8 / (time.perf_counter() - t0) / 1e9
    del dst

    # The call a user makes for a set of host-resident data sets: LevMarSolver.fit_host_batch (vp_fit_host_batch):
    # NTH worker threads of the library, one stream each, build -> fit -> read back one problem after the other.
    model_e2e = (vb.SeparableModelBuilder(["p0", "p1"]).function(["p0"], vb.ExpDecay()).function(["p1"], vb.ExpDecay())
                 .invariant_function(vb.Constant()).independent_variable(wls[0]["x"]).initial_parameters(alpha0).build())
    E2E_STEPS = int(os.environ.get("VP_E2E_STEPS", str(max(K, 60))))  # a longer region than K steps: ~40 ms instead of 14

    def e2e_run(nsteps):
        ys = [Yv[i % NTH] for i in range(nsteps)]
        reports, alpha, Cs = solver.fit_host_batch(model_e2e, ys, device=local_rank, workers=NTH)
        assert all(r.termination.was_successful() for r in reports)
        return alpha, Cs

    e2e_run(max(Wm, 2 * NTH))
    e2e_times = []
    for _ in range(3):
        barrier()
        t0 = time.perf_counter()
        alpha_e, Cs_e = e2e_run(E2E_STEPS)
        torch.cuda.synchronize()
        e2e_times.append(max_over_ranks(time.perf_counter() - t0) * K / E2E_STEPS)  # seconds per K steps
    for i in range(min(NTH, E2E_STEPS)):
        t = np.sort(np.asarray(wls[i % K]["alpha_true"], dtype=np.float64))
        assert np.max(np.abs(np.sort(alpha_e[i]) - t) / t) <= 1e-8
    clocks.__exit__(None, None, None)
    e2e_dt = _stats(e2e_times)["median"]
    e2e_val = world * K / e2e_dt

    # ---- N > 1: BASELINE config 5, both readings -----------------------------------------------------
    sharded = batch5 = None
    if world > 1:
        sharded = _sharded_global_fit(torch, dist, vb, api, solver, wls[0], rank, world, local_rank)
        batch5 = _independent_batch_c5(torch, dist, vb, rank, world, local_rank)

    # ---- roofline of the streaming kernel, extras, CPU baseline (rank 0) -------------------------------
    line = None
    if rank == 0:
        p = probs[0]
        pu, su = C.c_double(), C.c_double()
        g, sm = C.c_int64(), C.c_int64()
        lib.vp_profile_evaluation(p._h, 50, 0, C.byref(pu), C.byref(su), C.byref(g), C.byref(sm))
        warm_us = su.value
        lib.vp_profile_evaluation(p._h, 50, 512 << 20, C.byref(pu), C.byref(su), C.byref(g), C.byref(sm))
        cold_us = su.value
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        peak = float(peaks.get("hbm_gbs", 6650.0))
        achieved = timed_evals * BYTES_EVAL / (ms_med_local * 1e-3) / 1e9
        by_k["1"]["roofline_frac"] = (one[0][1][0] - 1) * BYTES_EVAL / (_stats([r[0] for r in one])["median"] * 1e-3) / 1e9 / peak
        if str(K_big) in by_k and K_big > K:
            by_k[str(K_big)]["roofline_frac"] = (by_k[str(K_big)]["timed_evaluations"] * BYTES_EVAL /
                                                  (by_k[str(K_big)]["ms"]["median"] * 1e-3) / 1e9 / peak)
        single_cold = BYTES_EVAL / (cold_us * 1e-6) / 1e9
        traffic_static = None
        try:
            traffic_static = json.load(open(os.path.join(ROOT, "profiles", "round2_queue_traffic.json")))
        except Exception:
            pass
        extras = {}
        if world == 1 and not args.quick:
            extras["c3"] = _extra_c3(torch, vb, lib, ctx, local_rank)
            extras["c4"] = _extra_c4(torch, vb, W, solver, peak)
        # CPU baseline: oracle port, 1 thread, bounded sample (rank 0, N = 1 only)
        cpu = None
        if world == 1 and not args.no_cpu:
            wl_cpu = c2_workload()
            NCPU = 10  # bounded sample: ~10 s of single-thread CPU work
            cruns = [cpu_fit_seconds(wl_cpu, 1) for _ in range(NCPU)]
            dtc = sum(r[0] for r in cruns)
            rep = cruns[-1][1]
            cpu = {"value": NCPU / dtc, "unit": "fits/s", "cores": 1, "kind": "port",
                   "sample": f"{NCPU} complete C2 fits (S={S_C2}, {rep['number_of_evaluations']} residual + "
                             f"{rep['number_of_jacobians']} Jacobian evaluations each, {dtc:.1f} s in total) with the C "
                             "restatement of varpro 0.13.3 (oracle/varpro_oracle.c), single thread like the reference; "
                             "not the Rust binary"}
        fits = world * K
        line = {
            "metric": METRIC, "value": fits / (ms_med * 1e-3), "unit": "fits/s",
            "n_gpus": world, "steps": K, "warmup": Wm, "ms_per_step": ms_med / K, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": "C2", "m": M, "S": S_C2, "n": N_BASIS, "q": Q, "alpha0": [2.0, 6.5],
                       "problems_per_rank": K, "concurrent_fits": K, "rank_ms": rank_ms,
                       "distinct_problems": "problem 0 = the canonical C2 instance; problems k > 0: own seed, tau* = (1, 3) * "
                                            "U[0.85, 1.15), own coefficients (noise-free like the reference bench)",
                       "timed_launches": R, "launch_ms": {"median": ms_med, "min": ms_min, "max": ms_max},
                       "value_from": "median launch, max over ranks",
                       "l2_policy": "K distinct 33.5 MB problems are fitted concurrently: K*33.5 MB of inputs are streamed "
                                    "per evaluation round (larger than the 126 MB L2 for K >= 4); K = 1: L2 flushed before every launch",
                       "evals_per_fit": nfev, "evals_per_fit_mean": float(np.mean(nfev)),
                       "fit_many_equals_fit": "same evaluation counts as the sequential vp_fit runs (asserted)"},
            "clocks": clocks.summary(),
            "e2e": {"value": e2e_val, "unit": "fits/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "host_threads": NTH, "timed_steps": E2E_STEPS, "api": "LevMarSolver.fit_host_batch -> vp_fit_host_batch",
                    "runs_s_per_K_steps": e2e_times,
                    "h2d_GBps_per_gpu": K * h2d / e2e_dt / 1e9,
                    "h2d_GBps_raw_memcpy_all_ranks_concurrent": raw_gbps,
                    "ceiling": "the PCIe link of each GPU (Gen5 x16): every step moves 33.6 MB host-to-device"},
            "gpu_launches": int(launches),
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": None if traffic_static is None else traffic_static.get("dram_bytes_per_launch"),
                         "traffic_source": None if traffic_static is None else
                         "static: ncu --set full capture of the same launch configuration, " + str(traffic_static.get("source")),
                         "kernel": "fit_queue_kernel<double,3,2,32,8> (all K fits on one persistent grid: work queue of "
                                   "(fit, item) work items; panel + Y-streaming reduce + LM step in-kernel)",
                         "how": "ONE launch fits all K problems (vp_fit_many): achieved = algorithmic bytes of the launch "
                                "(timed evaluations x 8*m*S) / CUDA-event time of the median timed launch on rank 0",
                         "bytes_per_launch": BYTES_EVAL * timed_evals, "timed_evaluations": timed_evals,
                         "fits_per_launch": K, "region_us": ms_med_local * 1e3,
                         "single_evaluation_full_grid": {"us_l2_flushed": cold_us, "GBps_l2_flushed": single_cold,
                                                         "frac_l2_flushed": single_cold / peak, "us_l2_warm": warm_us,
                                                         "GBps_l2_warm": BYTES_EVAL / (warm_us * 1e-6) / 1e9,
                                                         "grid": int(g.value), "smem_bytes": int(sm.value)},
                         "peak_source": "MEASURED_PEAKS.json hbm_gbs" if peaks else "fallback 6650 GB/s"},
            "by_concurrency": by_k,
            "latency_mode": {"value": world * K / (ms_seq * 1e-3), "unit": "fits/s", "ms_per_fit": ms_seq / K,
                             "what": "the same K fits one after the other (vp_fit), each on the whole GPU",
                             "evals_per_fit_mean": float(np.mean(nfev_seq))},
            "cpu_baseline": cpu,
        }
        if extras:
            line["extra"] = extras
        if sharded is not None:
            line["sharded_global_fit"] = sharded
        if batch5 is not None:
            line["independent_batch_c5"] = batch5
        print(json.dumps(line))
    for p in probs:
        p.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def _extra_c3(torch, vb, lib, ctx, device):
    """BASELINE config 3 at full size: 65 536 independent triple-exponential problems of 4096 samples, one
    vp_batch_fit launch. Bound: fp64 ALU (exp + Householder), against peaks measured in THIS run."""
    P, m = 65536, 4096
    dev = torch.device("cuda", device)
    gen = torch.Generator(device=dev)
    gen.manual_seed(65536)
    x = torch.linspace(0.0, 20.0, m, dtype=torch.float64, device=dev)
    tau = torch.tensor([1.0, 3.0, 9.0], dtype=torch.float64, device=dev) * (0.8 + 0.45 * torch.rand(P, 3, generator=gen, device=dev, dtype=torch.float64))
    c = 1.0 + 9.0 * torch.rand(P, 3, generator=gen, device=dev, dtype=torch.float64)
    Y = torch.zeros(P, m, dtype=torch.float64, device=dev)  # (P, m) row-major == m x P column-major
    for j in range(3):
        Y += c[:, j:j + 1] * torch.exp(-x[None, :] / tau[:, j:j + 1])
    Y += 1e-3 * torch.randn(P, m, generator=gen, device=dev, dtype=torch.float64)
    alpha0 = (tau * torch.tensor([1.3, 0.8, 1.2], dtype=torch.float64, device=dev)).cpu().numpy()
    torch.cuda.synchronize()
    model = (vb.SeparableModelBuilder(["t1", "t2", "t3"]).function(["t1"], vb.ExpDecay()).function(["t2"], vb.ExpDecay())
             .function(["t3"], vb.ExpDecay()).independent_variable(x.cpu().numpy()).initial_parameters([1.0, 3.0, 9.0]).build())
    batch = vb.IndependentBatch(model, None, alpha0, y_device_ptr=Y.data_ptr(), P=P)
    del Y
    ext = torch.cuda.ExternalStream(ctx.stream(), device=device)
    times = []
    for it in range(4):
        batch.set_params(alpha0)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        with torch.cuda.stream(ext):
            e0.record()
        batch.fit(reports=False)
        with torch.cuda.stream(ext):
            e1.record()
        torch.cuda.synchronize()
        if it > 0:
            times.append(e0.elapsed_time(e1))
    batch.set_params(alpha0)
    res = batch.fit()
    nfev = res.number_of_evaluations
    batch.close()
    fma, ex = C.c_double(), C.c_double()
    lib.vp_measure_fp64_peaks(ctx.h, C.byref(fma), C.byref(ex))
    ms = _stats(times)["median"]
    evals = float(nfev.sum())
    # algorithmic work of one evaluation: n*m fp64 exp + ~2*m*(n+p+1)*(n+2) fused multiply-adds of the
    # Householder steps, projections and the tail reduction
    n, p = 3, 3
    exps = n * m
    flops = 2.0 * m * (n + p + 1) * (n + 2) * 2
    floor_s = evals * (exps / (ex.value * 1e9) + flops / (fma.value * 1e12))
    return {"workload": "C3: triple-exponential decay, 4096 samples, 65 536 independent problems, fp64, one vp_batch_fit launch",
            "fits_per_s": P / (ms * 1e-3), "ms_per_batch": _stats(times), "evaluations_mean": float(nfev.mean()),
            "evaluations_max": int(nfev.max()), "converged_fraction": float(res.successful.mean()),
            "roofline": {"bound": "fp64 ALU (exp + DFMA)", "achieved": floor_s / (ms * 1e-3), "peak": 1.0, "unit": "fraction of the ALU floor",
                         "frac": floor_s / (ms * 1e-3), "dfma_tflops_measured": fma.value, "dexp_gexps_measured": ex.value,
                         "alu_floor_ms": 1e3 * floor_s, "hbm_floor_ms": 1e3 * 8.0 * m * P / 6.5e12,
                         "how": "evaluations made x (3*4096 exp / measured exp rate + ~0.29 MFLOP / measured DFMA rate) / launch time; "
                                "peaks measured by vp_measure_fp64_peaks in this run, same clocks record"}}


def _extra_c4(torch, vb, W, solver, peak):
    """BASELINE config 4 at full size: weighted multi-exponential (reference lmfit asset shape), m = 1000, S = 16 384,
    fp32 in HBM / fp64 arithmetic: one global fit + fit statistics of every column."""
    wl = W.c4()
    gp = W.make_gpu_problem(wl, dtype=np.float32)
    ctx = gp._ctx
    ext = torch.cuda.ExternalStream(ctx.stream())
    times, nf = [], 0
    for it in range(8):
        gp.set_params(wl["alpha0"])
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        with torch.cuda.stream(ext):
            e0.record()
        res = solver.fit(gp)
        with torch.cuda.stream(ext):
            e1.record()
        torch.cuda.synchronize()
        nf = res.minimization_report.number_of_evaluations
        if it > 1:
            times.append(e0.elapsed_time(e1))
    t0 = time.perf_counter()
    st = gp.statistics()
    t_stat = time.perf_counter() - t0
    chi2 = st[0].reduced_chi2()
    gp.close()
    ms = _stats(times)["median"]
    bytes_eval = 4 * 1000 * 16384
    ach = (nf - 1) * bytes_eval / (ms * 1e-3) / 1e9
    return {"workload": "C4: weighted multiexp (test_assets/weighted_multiexp_decay shape), m=1000, MRHS=16384, fp32 in HBM, "
                        "fp64 arithmetic, global fit + per-column FitStatistics",
            "fits_per_s": 1e3 / ms, "ms_per_fit": _stats(times), "evaluations": nf,
            "statistics_ms_all_columns_incl_d2h": 1e3 * t_stat, "reduced_chi2_column0": chi2,
            "roofline": {"bound": "hbm", "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak,
                         "how": "(evaluations - 1) x 4*m*S bytes / CUDA-event time of one vp_fit (persistent whole-fit kernel fit_kernel_dmma<float>: fp32 tiles converted to fp64 at the fragment loads; the time includes the launch and the read-back of the fit state)"}}


def _sharded_global_fit(torch, dist, vb, api, solver, wl, rank, world, device, cols_per_rank=131072, reps=5):
    """BASELINE config 5: double-exponential MRHS with 131 072 columns per GPU (1.07 GB fp64 each), ONE global
    fit whose per-GPU sums are exchanged through the NVLink mailboxes inside the fit kernel. Checked without the
    oracle (8.6 GB): recovered parameters against the truth, every coefficient against the generating C*."""
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
    alpha = r.nonlinear_parameters()
    assert np.allclose(np.sort(alpha), [1.0, 3.0], atol=1e-8), alpha
    Cg = torch.from_numpy(np.ascontiguousarray(r.linear_coefficients())).cuda()
    if alpha[0] > alpha[1]:
        Cg = Cg[[1, 0, 2]]
    cerr = torch.tensor([float((Cg - Cs).abs().max())], dtype=torch.float64, device="cuda")
    rn = torch.tensor([np.sqrt(2 * r.minimization_report.objective_function)], dtype=torch.float64, device="cuda")
    dist.all_reduce(cerr, op=dist.ReduceOp.MAX)
    rn_all = [torch.zeros_like(rn) for _ in range(world)]
    dist.all_gather(rn_all, rn)
    assert len({float(v.item()) for v in rn_all}) == 1, "all ranks must hold the same global residual norm"
    assert float(cerr.item()) <= 1e-6, float(cerr.item())
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
            "checks": {"max_abs_coefficient_error_vs_generating_C": float(cerr.item()), "alpha": [float(a) for a in alpha],
                       "global_residual_norm": float(rn.item()), "ranks_agree_bitwise_on_residual_norm": True},
            "collective": "in-kernel NVLink mailbox exchange of the per-GPU sums (||r||^2, G, V, U), one per evaluation; no NCCL on the data path"}


def _independent_batch_c5(torch, dist, vb, rank, world, device, per_rank=131072):
    """BASELINE config 5's other reading: 1 048 576 INDEPENDENT double-exponential problems (1024 samples each)
    over the GPUs through vp_batch_*: partitioned, no collective, one launch per GPU."""
    x = np.linspace(0.0, 12.5, M)
    dev = torch.device("cuda", device)
    gen = torch.Generator(device=dev)
    gen.manual_seed(5000 + rank)
    xd = torch.from_numpy(x).to(dev)
    tau = torch.tensor([1.0, 3.0], dtype=torch.float64, device=dev) * (0.85 + 0.3 * torch.rand(per_rank, 2, generator=gen, device=dev, dtype=torch.float64))
    c = 1.0 + 9.0 * torch.rand(per_rank, 3, generator=gen, device=dev, dtype=torch.float64)
    Y = c[:, 2:3].expand(per_rank, M).clone()
    for j in range(2):
        Y += c[:, j:j + 1] * torch.exp(-xd[None, :] / tau[:, j:j + 1])
    Y += 1e-4 * torch.randn(per_rank, M, generator=gen, device=dev, dtype=torch.float64)
    alpha0 = (tau * torch.tensor([1.25, 0.8], dtype=torch.float64, device=dev)).cpu().numpy()
    model = (vb.SeparableModelBuilder(["p0", "p1"]).function(["p0"], vb.ExpDecay()).function(["p1"], vb.ExpDecay())
             .invariant_function(vb.Constant()).independent_variable(x).initial_parameters([1.0, 3.0]).build())
    batch = vb.IndependentBatch(model, None, alpha0, y_device_ptr=Y.data_ptr(), P=per_rank, device=device)
    del Y
    times = []
    for it in range(4):
        batch.set_params(alpha0)
        torch.cuda.synchronize()
        dist.barrier()
        t0 = time.perf_counter()
        batch.fit(reports=False)
        torch.cuda.synchronize()
        if it > 0:
            times.append(time.perf_counter() - t0)
    batch.set_params(alpha0)
    res = batch.fit()
    tau_h = tau.cpu().numpy()
    err = np.max(np.abs(np.sort(res.nonlinear_parameters, axis=1) - np.sort(tau_h, axis=1)) / np.sort(tau_h, axis=1), axis=1)
    ok = torch.tensor([float(res.successful.mean()), float((err < 1e-2).mean())], dtype=torch.float64, device="cuda")
    dist.all_reduce(ok, op=dist.ReduceOp.SUM)
    t = torch.tensor([_stats(times)["median"]], dtype=torch.float64, device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    batch.close()
    return {"workload": "C5 as independent problems: double-exponential + offset, 1024 samples, one single-RHS problem per column, "
                        "partitioned over the ranks (vp_batch_*), no collective",
            "problems_total": per_rank * world, "problems_per_gpu": per_rank, "fits_per_s": per_rank * world / float(t.item()),
            "ms_per_batch_max_over_ranks": 1e3 * float(t.item()), "converged_fraction": float(ok[0].item()) / world,
            "recovered_within_1pct_fraction": float(ok[1].item()) / world, "evaluations_mean": float(res.number_of_evaluations.mean())}


def _build_on_device(wl, device):
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
    ap.add_argument("--repeats", type=int, default=10, help="timed launches of the K-fit batch (value = median)")
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--quick", action="store_true", help="skip the K = 60 leg and the config 3 / config 4 extras")
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
