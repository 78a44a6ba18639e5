#!/bin/bash
# Final evidence visit of a round (one GPU): whole GPU suite, smoke, bench (+ reference arm) with a clocks record,
# launch list of the bench command, full ncu captures of the three hot kernels, racecheck of the batch kernel.
# Run as: gpurun --timeout 2700 -- 'bash scripts/gpu_final.sh <tag>'
TAG=${1:-final}
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap --format=csv -lms 500 > gpurun_out/${TAG}_clocks.csv &
SMI=$!
nproc > gpurun_out/${TAG}_host.txt; lscpu | grep -E "Model name|^CPU\(s\)" >> gpurun_out/${TAG}_host.txt
( time timeout 1500 python -m pytest tests -m gpu -q --durations=8 ) > gpurun_out/${TAG}_pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/${TAG}_pytest_gpu.log
tail -6 gpurun_out/${TAG}_pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${TAG}_smoke.log 2>&1; tail -2 gpurun_out/${TAG}_smoke.log
( time timeout 900 python bench.py --steps 20 --warmup 5 ) > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; tail -4 gpurun_out/${TAG}_bench.err
timeout 400 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${TAG}_bench_ref.json 2>> gpurun_out/${TAG}_bench.err; cut -c1-400 gpurun_out/${TAG}_bench_ref.json
kill $SMI
for m in 4096 1000 256; do VP_BATCH_DBG=1 timeout 120 python scripts/bench_c3.py 65536 $m 2>&1 | tail -2; done > gpurun_out/${TAG}_batch_phases.txt
VP_QUEUE_DBG=1 timeout 200 python bench.py --steps 20 --warmup 3 --repeats 5 --no-cpu --quick 2>&1 | grep "queue dbg" | tail -1 > gpurun_out/${TAG}_queue_phases.txt
FIT_TIMELINE=1 timeout 200 python scripts/gpu_probe.py > gpurun_out/${TAG}_fit_timeline.txt 2>&1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/${TAG}_launches.csv python bench.py --steps 20 --warmup 3 --repeats 3 --no-cpu --quick > gpurun_out/${TAG}_ncu_bench.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:fit_queue -s 3 -c 1 -f -o gpurun_out/${TAG}_queue python bench.py --steps 20 --warmup 3 --repeats 3 --no-cpu --quick > gpurun_out/${TAG}_ncu_queue.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:fit_kernel_dmma -s 8 -c 1 -f -o gpurun_out/${TAG}_fit python scripts/ncu_fit.py > gpurun_out/${TAG}_ncu_fit.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:batch_fit -c 1 -f -o gpurun_out/${TAG}_batch python scripts/bench_c3.py 16384 > gpurun_out/${TAG}_ncu_batch.log 2>&1
timeout 900 compute-sanitizer --tool racecheck python -m pytest tests -m gpu -x -q -k "batch_slots_refill and 200" > gpurun_out/${TAG}_racecheck_batch.txt 2>&1; tail -5 gpurun_out/${TAG}_racecheck_batch.txt
timeout 900 compute-sanitizer --tool memcheck python -m pytest tests -m gpu -x -q -k "batch_slots_refill or further_fast_path_shapes_independent" > gpurun_out/${TAG}_memcheck_batch.txt 2>&1; tail -5 gpurun_out/${TAG}_memcheck_batch.txt
ls -la gpurun_out/${TAG}_* | tail -25
