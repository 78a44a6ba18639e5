"""BASELINE config 3 on the GPU: triple-exponential decay, m = 4096 samples, P independent problems
(default 65 536 = 2.1 GB of observations), fp64, one vp_batch_fit launch. Prints one JSON line.
Data are generated on the device (torch is the harness' data generator only)."""
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import varpro_b200 as vb  # noqa: E402

P = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
m = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
dev = torch.device("cuda:0")
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
times = []
for it in range(3):
    batch.set_params(alpha0)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    batch.fit(reports=False)
    torch.cuda.synchronize()
    times.append(time.perf_counter() - t0)
batch.set_params(alpha0)
res = batch.fit()
tau_h = tau.cpu().numpy()
ok = res.successful
err = np.max(np.abs(np.sort(res.nonlinear_parameters, axis=1) - np.sort(tau_h, axis=1)) / np.sort(tau_h, axis=1), axis=1)
best = min(times)
nfev = res.number_of_evaluations
print(json.dumps({"workload": "C3: triple-exponential, independent batch", "m": m, "P": P, "dtype": "f64",
                  "fits_per_s": P / best, "ms_per_batch": 1e3 * best, "evaluations_mean": float(nfev.mean()),
                  "evaluations_max": int(nfev.max()), "converged_fraction": float(ok.mean()),
                  "recovered_within_5pct_fraction": float((err < 0.05).mean()),
                  "us_per_evaluation_per_sm": 1e6 * best * 148 / float(nfev.sum()),
                  "hbm_GBps_of_y": 8.0 * m * P / best / 1e9}))
