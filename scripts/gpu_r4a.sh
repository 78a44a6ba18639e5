#!/bin/bash
# batch kernel iteration: batch parity tests + C3 throughput (+ optional ncu of the batch kernel)
TAG=${1:-r4a}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -k "batch or c3" > gpurun_out/${TAG}_pytest_batch.log 2>&1; echo "pytest exit $?" >> gpurun_out/${TAG}_pytest_batch.log
tail -15 gpurun_out/${TAG}_pytest_batch.log
timeout 600 python scripts/bench_c3.py 65536 4096 2>&1 | tee gpurun_out/${TAG}_c3.json
timeout 300 python scripts/bench_c3.py 65536 1000 2>&1 | tee gpurun_out/${TAG}_c3_m1000.json
if [ "$2" = "ncu" ]; then
timeout 600 ncu --set full --import-source on --clock-control none -k regex:batch_fit_kernel -c 1 -o gpurun_out/${TAG}_batch -f python scripts/bench_c3.py 8192 4096 > gpurun_out/${TAG}_ncu_batch.log 2>&1
fi
