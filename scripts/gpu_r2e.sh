#!/bin/bash
# Round-2 GPU visit E: the whole GPU test-suite (incl. config 3 at full size), smoke, the full bench line.
TAG=${1:-r2e}
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -q --durations=8 ) > gpurun_out/${TAG}_pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/${TAG}_pytest_gpu.log
tail -40 gpurun_out/${TAG}_pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${TAG}_smoke.log 2>&1; tail -3 gpurun_out/${TAG}_smoke.log
( time timeout 900 python bench.py --steps 20 --warmup 3 ) > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; tail -8 gpurun_out/${TAG}_bench.err
python - <<PY
import json
d=json.loads(open("gpurun_out/%s_bench.json" % "${TAG}").read().splitlines()[0])
print("value",round(d["value"]),"frac",round(d["roofline"]["frac"],3),"evals",d["config"]["evals_per_fit_mean"],"launch_ms",d["config"]["launch_ms"])
print("by_k", {k:(round(v["fits_per_s"]), v.get("roofline_frac")) for k,v in d["by_concurrency"].items()})
print("latency",round(d["latency_mode"]["value"]),"e2e",d["e2e"])
print("c3", d["extra"]["c3"]["fits_per_s"], d["extra"]["c3"]["roofline"]["frac"], "c4", d["extra"]["c4"]["fits_per_s"], d["extra"]["c4"]["roofline"]["frac"])
print("cpu", d["cpu_baseline"]["value"])
PY
