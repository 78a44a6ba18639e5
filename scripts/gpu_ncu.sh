#!/bin/bash
# ncu evidence for the bench command: launch list + one full capture of the timed work-queue launch (< 64 MiB)
TAG=${1:-r01}
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_launches.csv python bench.py --steps 20 --warmup 3 --no-cpu > gpurun_out/${TAG}_ncu_bench.log 2>&1
# fit_queue_kernel launches of the bench: 3 warm-up batches, then the timed one
timeout 900 ncu --set full --clock-control none --import-source on -k regex:fit_queue -s 3 -c 1 -f -o gpurun_out/${TAG}_queue python bench.py --steps 20 --warmup 3 --no-cpu > gpurun_out/${TAG}_ncu_full.log 2>&1
ls -la gpurun_out | tail -5
