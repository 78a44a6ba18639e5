#!/bin/bash
# Round-2 GPU visit B: new bench.py (distinct problems, repeated launches, K = 1 / 20 / 60, C3 / C4 extras), smoke,
# 8- vs 16-warp tilings, items-per-CTA sweep with the phase breakdown.
TAG=${1:-r2b}
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -x -k "bitwise or fit_many or world2 or fit_modes or fused_kernel or c2_full_size" > gpurun_out/${TAG}_pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/${TAG}_pytest_gpu.log
tail -8 gpurun_out/${TAG}_pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${TAG}_smoke.log 2>&1; tail -3 gpurun_out/${TAG}_smoke.log
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; tail -5 gpurun_out/${TAG}_bench.err; cat gpurun_out/${TAG}_bench.json
summ='import json,sys
d=json.loads(sys.stdin.read()); print("value",round(d["value"]),"frac",round(d["roofline"]["frac"],3),"evals",d["config"]["evals_per_fit_mean"],"launch_ms",d["config"]["launch_ms"],"K1",round(d["by_concurrency"]["1"]["fits_per_s"]),"latency",round(d["latency_mode"]["value"]),"e2e",round(d["e2e"]["value"]), "raw_h2d", round(d["e2e"]["h2d_GBps_raw_memcpy_all_ranks_concurrent"],1))'
for w in 8 16; do
for ipc in 1 2 3; do
  echo "== warps=$w K=20 items_per_cta=$ipc" >> gpurun_out/${TAG}_sweep.txt
  VP_FIT_WARPS=$w VP_QUEUE_ITEMS_PER_CTA=$ipc VP_QUEUE_DBG=1 timeout 300 python bench.py --steps 20 --warmup 3 --repeats 5 --no-cpu --quick 2>> gpurun_out/${TAG}_sweep.txt | python -c "$summ" >> gpurun_out/${TAG}_sweep.txt
done
done
grep -E "^==|^value|queue dbg" gpurun_out/${TAG}_sweep.txt | awk '/queue dbg/{l=$0} /^==/{print} /^value/{print l; print}' | cut -c1-640
VP_FIT_WARPS=16 FIT_TIMELINE=1 timeout 300 python scripts/gpu_probe.py > gpurun_out/${TAG}_probe_w16.txt 2>&1; cat gpurun_out/${TAG}_probe_w16.txt
