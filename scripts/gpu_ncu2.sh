#!/bin/bash
# ncu evidence, round 2: launch list of the bench command; full captures of the timed work-queue launch, of one
# persistent whole-fit launch and of the independent-batch kernel.
TAG=${1:-r3a}
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/${TAG}_launches.csv python bench.py --steps 20 --warmup 3 --repeats 3 --no-cpu --quick > gpurun_out/${TAG}_ncu_bench.log 2>&1
# fit_queue_kernel launches of the bench: 3 warm-up batches, then the timed ones
timeout 900 ncu --set full --clock-control none --import-source on -k regex:fit_queue -s 3 -c 1 -f -o gpurun_out/${TAG}_queue python bench.py --steps 20 --warmup 3 --repeats 3 --no-cpu --quick > gpurun_out/${TAG}_ncu_queue.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:fit_kernel_dmma -s 8 -c 1 -f -o gpurun_out/${TAG}_fit python scripts/ncu_fit.py > gpurun_out/${TAG}_ncu_fit.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:batch_fit -s 1 -c 1 -f -o gpurun_out/${TAG}_batch python scripts/bench_c3.py 8192 > gpurun_out/${TAG}_ncu_batch.log 2>&1
ls -la gpurun_out/${TAG}_* | tail -8
