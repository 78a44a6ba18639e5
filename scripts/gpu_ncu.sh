#!/bin/bash
# ncu evidence for the bench command: launch list + one full capture of a whole-fit launch of the
# persistent kernel (kept < 64 MiB) + one of a single-evaluation launch.
TAG=${1:-r01}
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_launches.csv python bench.py --steps 20 --warmup 3 --no-cpu > gpurun_out/${TAG}_ncu_bench.log 2>&1
# bench launch order per warm-up batch: 20 single-evaluation launches (problem builds), then 20 whole-fit launches
timeout 900 ncu --set full --clock-control none --import-source on -k regex:fit_kernel -s 25 -c 1 -f -o gpurun_out/${TAG}_fit python bench.py --steps 20 --warmup 3 --no-cpu > gpurun_out/${TAG}_ncu_full.log 2>&1
ls -la gpurun_out
