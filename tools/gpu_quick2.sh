#!/bin/bash
# GPU tests + bench lines for the named workloads.  TAG=$1  WORKLOADS=$2
TAG=${1:-q}
WLS=${2:-"cfg2 cfg4 cfg5"}
mkdir -p gpurun_out
timeout -s KILL 900 python -m pytest tests -m gpu -q 2>&1 | tail -30 | tee gpurun_out/pytest_${TAG}.log
for wl in $WLS; do
  timeout -s KILL 600 python bench.py --workload $wl --steps 20 --warmup 5 --no-cpu > gpurun_out/bench_${wl}_${TAG}.json 2> gpurun_out/bench_${wl}_${TAG}.err || tail -8 gpurun_out/bench_${wl}_${TAG}.err
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/bench_${wl}_${TAG}.json").read().strip().splitlines()[-1])
    print("${wl}", round(d["value"],1), "audio-h/s", round(d["ms_per_step"],4), "ms", d.get("stages_ms"), "pipeline_frac", round(d["roofline"].get("pipeline_frac"),3), "mem", d["config"].get("device_memory_gb"), "e2e", {k:d.get("e2e",{}).get(k) for k in ("value","h2d_ceiling_gbs","h2d_achieved_gbs","frac_of_h2d_ceiling")}, "launches", d.get("gpu_launches"))
except Exception as e: print("${wl} failed", e)
PY
done
