#!/bin/bash
# quick GPU visit: parity tests
TAG=${1:-q}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/${TAG}_pytest_gpu.log
tail -40 gpurun_out/${TAG}_pytest_gpu.log
