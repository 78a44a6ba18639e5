#!/bin/bash
# full GPU suite + batch phases + single-fit timeline
TAG=${1:-r4d}
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/${TAG}_pytest_gpu.log
tail -15 gpurun_out/${TAG}_pytest_gpu.log
for m in 4096 1000 256; do VP_BATCH_DBG=1 timeout 120 python scripts/bench_c3.py 65536 $m 2>&1 | tail -2; done | tee gpurun_out/${TAG}_c3.txt
timeout 300 python scripts/trace_fit.py 2>&1 | tee gpurun_out/${TAG}_trace_fit.txt | tail -25
