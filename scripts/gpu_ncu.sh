#!/bin/bash
# ncu evidence: launch list of the bench command + one full capture of the streaming kernel (kept < 64 MiB).
TAG=${1:-r01}
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/${TAG}_ncu_bench.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:stream_kernel -s 30 -c 1 -f -o gpurun_out/${TAG}_k2 python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/${TAG}_ncu_full.log 2>&1
S=4096 TIMELINE=1 timeout 300 python scripts/gpu_probe.py > gpurun_out/${TAG}_probe_4096.log 2>&1
S=131072 timeout 300 python scripts/gpu_probe.py > gpurun_out/${TAG}_probe_131072.log 2>&1
VP_TRACE=1 S=4096 timeout 300 python scripts/gpu_probe.py 2>&1 | grep "vp_fit" | head -60 > gpurun_out/${TAG}_trace.log
ls -la gpurun_out
