#!/bin/bash
# quick GPU visit: parity tests + bench
TAG=${1:-q}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/${TAG}_pytest_gpu.log
tail -25 gpurun_out/${TAG}_pytest_gpu.log
timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; cat gpurun_out/${TAG}_bench.json; tail -5 gpurun_out/${TAG}_bench.err
timeout 600 python bench.py --steps 60 --warmup 3 --no-cpu > gpurun_out/${TAG}_bench60.json 2>> gpurun_out/${TAG}_bench.err; cat gpurun_out/${TAG}_bench60.json; tail -5 gpurun_out/${TAG}_bench.err
