#!/bin/bash
# Round-2 GPU visit C: adaptive work-item sizing; determinism on the 16-warp tiling; K sweep with phase breakdown;
# e2e host-thread sweep.
TAG=${1:-r2c}
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -x -k "bitwise or deterministic or fit_many or reusable or world2 or c_program" > gpurun_out/${TAG}_pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/${TAG}_pytest_gpu.log
tail -8 gpurun_out/${TAG}_pytest_gpu.log
summ='import json,sys
d=json.loads(sys.stdin.read()); print("value",round(d["value"]),"frac",round(d["roofline"]["frac"],3),"evals",d["config"]["evals_per_fit_mean"],"launch_ms",d["config"]["launch_ms"],"K1",round(d["by_concurrency"]["1"]["fits_per_s"]),"latency",round(d["latency_mode"]["value"]),"e2e",round(d["e2e"]["value"]), "raw_h2d", round(d["e2e"]["h2d_GBps_raw_memcpy_all_ranks_concurrent"],1))'
for K in 20 60 8; do
for ipc in 1 2 4; do
  echo "== K=$K items_per_cta=$ipc" >> gpurun_out/${TAG}_sweep.txt
  VP_QUEUE_ITEMS_PER_CTA=$ipc VP_QUEUE_DBG=1 timeout 300 python bench.py --steps $K --warmup 3 --repeats 5 --no-cpu --quick 2>> gpurun_out/${TAG}_sweep.txt | python -c "$summ" >> gpurun_out/${TAG}_sweep.txt
done
done
grep -E "^==|^value|queue dbg" gpurun_out/${TAG}_sweep.txt | awk '/queue dbg/{l=$0} /^==/{print} /^value/{print l; print}' | cut -c1-640
for th in 2 4 6; do
  echo "== e2e threads=$th" >> gpurun_out/${TAG}_e2e.txt
  VP_E2E_THREADS=$th timeout 300 python bench.py --steps 20 --warmup 3 --repeats 3 --no-cpu --quick 2>> gpurun_out/${TAG}_e2e.txt | python -c "$summ" >> gpurun_out/${TAG}_e2e.txt
done
cat gpurun_out/${TAG}_e2e.txt | grep -E "^==|^value"
