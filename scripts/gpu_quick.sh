#!/bin/bash
# quick GPU visit: parity tests + bench
TAG=${1:-q}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/${TAG}_pytest_gpu.log
tail -30 gpurun_out/${TAG}_pytest_gpu.log
for K in 20 60; do
timeout 600 python bench.py --steps $K --warmup 3 --no-cpu > gpurun_out/${TAG}_bench$K.json 2> gpurun_out/${TAG}_bench.err; tail -3 gpurun_out/${TAG}_bench.err
python -c "import sys,json; d=json.loads(open('gpurun_out/${TAG}_bench$K.json').read()); print('K=$K value', round(d['value']), 'e2e', round(d['e2e']['value']), 'frac', round(d['roofline']['frac'],3), 'lat', round(d['latency_mode']['value']), 'evals', d['config']['evals_per_fit_mean'])"
done
timeout 300 python scripts/bench_c4.py f32 > gpurun_out/${TAG}_c4_f32.json 2>&1; cat gpurun_out/${TAG}_c4_f32.json
