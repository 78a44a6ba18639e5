#!/bin/bash
# Round-2 GPU visit A: parity of the rewritten kernels, phase breakdowns, items-per-CTA sweep.
# Run as: gpurun --timeout 1500 -- 'bash scripts/gpu_r2a.sh <tag>'
TAG=${1:-r2a}
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -x --deselect tests/test_gpu_round2.py::test_c3_full_size_sampled_against_the_oracle > gpurun_out/${TAG}_pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/${TAG}_pytest_gpu.log
tail -30 gpurun_out/${TAG}_pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${TAG}_smoke.log 2>&1; tail -3 gpurun_out/${TAG}_smoke.log
FIT_TIMELINE=1 timeout 300 python scripts/gpu_probe.py > gpurun_out/${TAG}_probe.txt 2>&1; cat gpurun_out/${TAG}_probe.txt
for ipc in 1 2 3 4; do
  for K in 20; do
    echo "== K=$K items_per_cta=$ipc" >> gpurun_out/${TAG}_sweep.txt
    VP_QUEUE_ITEMS_PER_CTA=$ipc VP_QUEUE_DBG=1 timeout 300 python bench.py --steps $K --warmup 3 --no-cpu 2>> gpurun_out/${TAG}_sweep.txt | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('value',round(d['value']),'frac',round(d['roofline']['frac'],3),'evals',d['config']['evals_per_fit_mean'],'latency',round(d['latency_mode']['value']),'e2e',round(d['e2e']['value']))" >> gpurun_out/${TAG}_sweep.txt
  done
done
grep -E "^==|^value|queue dbg" gpurun_out/${TAG}_sweep.txt | awk '/queue dbg/{l=$0} /^==/{print} /^value/{print l; print}' | cut -c1-600
