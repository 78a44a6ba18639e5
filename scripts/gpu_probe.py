"""Quick GPU probe: C2 evaluation timings (host-driven), fit timing, per-kernel event times."""
import ctypes as C
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import workloads as W  # noqa: E402
import varpro_b200 as vb  # noqa: E402
from varpro_b200 import _lib  # noqa: E402


def prof(gp, iters=20, flush=0):
    p, s = C.c_double(), C.c_double()
    g, sm = C.c_int64(), C.c_int64()
    st = _lib.load().vp_profile_evaluation(gp._h, iters, flush, C.byref(p), C.byref(s), C.byref(g), C.byref(sm))
    assert st == 0, st
    return p.value, s.value, g.value, sm.value


S = int(os.environ.get("S", 4096))
wl = W.c2(S=S)
t0 = time.time()
gp = W.make_gpu_problem(wl)
print("create: %.1f ms" % (1e3 * (time.time() - t0)))
bytes_eval = 8 * 1024 * S
for flush in (0, 256 << 20):
    prof(gp, 3, flush)
    p, s, g, sm = prof(gp, 20, flush)
    print(f"flush={flush>>20}MB panel {p:.2f} us  stream {s:.2f} us  grid {g} smem {sm}  "
          f"=> {bytes_eval / s / 1e3:.1f} GB/s ({bytes_eval / s / 1e3 / 6552.6:.3f} of measured HBM peak)")
t0 = time.time()
res = vb.LevMarSolver.default().fit(gp)
dt = time.time() - t0
rep = res.minimization_report
print(f"fit: {1e3*dt:.2f} ms, nfev {rep.number_of_evaluations}, term {rep.termination}, alpha {res.nonlinear_parameters()}")
for _ in range(3):
    gp.set_params(wl["alpha0"])
    t0 = time.time()
    res = vb.LevMarSolver.default().fit(gp)
    dt = time.time() - t0
    print(f"fit again: {1e3*dt:.2f} ms nfev {res.minimization_report.number_of_evaluations} -> {1e6*dt/res.minimization_report.number_of_evaluations:.1f} us/eval")

def dump_timeline(title):
    cap = (148 * 8 + 1) * 16
    buf = (C.c_longlong * cap)()
    g = C.c_int64()
    st = _lib.load().vp_debug_timeline(gp._h, buf, cap, C.byref(g))
    assert st == 0
    t = np.array(buf[: (g.value + 1) * 16], dtype=np.int64).reshape(-1, 16)
    k2 = t[:-1]
    print(title)
    names = {8: "eval_start", 9: "basis_evaluated", 10: "panel_done", 1: "frags_loaded", 2: "first_tile", 3: "loop_done",
             4: "publish_begin", 14: "counted_in", 5: "all_arrived", 12: "rows_folded", 13: "assembled",
             15: "lm_step_done(t0)", 6: "evaluation_end"}
    for i, nm in names.items():
        col = k2[:, i][k2[:, i] >= 0]
        if len(col):
            print(f"  {nm:18s} min {col.min():7d} median {int(np.median(col)):7d} max {col.max():7d} (n={len(col)})")


if os.environ.get("FIT_TIMELINE"):
    vb.set_option("dbg_fit", 1)
    gp.set_params(wl["alpha0"])
    res = vb.LevMarSolver.default().fit(gp)
    dump_timeline("persistent fit, last evaluation (ns relative to the earliest stamp):")

if os.environ.get("TIMELINE"):
    import numpy as np
    cap = (148 * 8 + 1) * 16
    buf = (C.c_longlong * cap)()
    g = C.c_int64()
    st = _lib.load().vp_debug_timeline(gp._h, buf, cap, C.byref(g))
    assert st == 0
    t = np.array(buf[: (g.value + 1) * 16], dtype=np.int64).reshape(-1, 16)
    k2, k1 = t[:-1], t[-1]
    print("K1 panel marks (ns):", k1[:5])
    names = {0: "start", 8: "bar_init_done", 9: "before_panel_wait", 10: "panel_landed", 1: "frags_loaded", 2: "first_tile", 3: "loop_done", 4: "publish_begin", 14: "partial_written", 5: "published", 11: "fin_prefetch", 12: "fin_partials", 13: "fin_assembled", 15: "lm_step_done", 6: "finalize_done"}
    for i, nm in names.items():
        col = k2[:, i][k2[:, i] >= 0]
        if len(col):
            print(f"K2 {nm:14s} min {col.min():7d} median {int(np.median(col)):7d} max {col.max():7d} (n={len(col)})")
