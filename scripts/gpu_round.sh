#!/bin/bash
# One GPU-box visit: parity tests, smoke, bench (+ reference arm), ncu launch list, one full capture of the
# timed work-queue launch, stress test.
# Run as: gpurun --timeout 2400 -- 'bash scripts/gpu_round.sh <tag>'
TAG=${1:-r01}
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap --format=csv -lms 500 > gpurun_out/${TAG}_clocks.csv &
SMI=$!
nproc > gpurun_out/${TAG}_host.txt; lscpu | grep -E "Model name|^CPU\(s\)" >> gpurun_out/${TAG}_host.txt
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/${TAG}_pytest_gpu.log
tail -5 gpurun_out/${TAG}_pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${TAG}_smoke.log 2>&1; tail -2 gpurun_out/${TAG}_smoke.log
timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; cat gpurun_out/${TAG}_bench.json; tail -3 gpurun_out/${TAG}_bench.err
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${TAG}_bench_ref.json 2>> gpurun_out/${TAG}_bench.err; cat gpurun_out/${TAG}_bench_ref.json
timeout 600 python scripts/stress_fit_many.py 1 30 > gpurun_out/${TAG}_stress.log 2>&1; tail -2 gpurun_out/${TAG}_stress.log
kill $SMI
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_launches.csv python bench.py --steps 20 --warmup 3 --no-cpu > gpurun_out/${TAG}_ncu_bench.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:fit_queue -s 3 -c 1 -f -o gpurun_out/${TAG}_queue python bench.py --steps 20 --warmup 3 --no-cpu > gpurun_out/${TAG}_ncu_full.log 2>&1
ls -la gpurun_out | tail -12
