#!/bin/bash
# short GPU call: GPU tests (no -x), cfg2 bench, launch list, full ncu of the named kernels.  TAG=$1  KERNELS=$2
TAG=${1:-q}
KFULL=${2:-logmel_tc_kernel}
mkdir -p gpurun_out
timeout -s KILL 900 python -m pytest tests -m gpu -q 2>&1 | tail -25 | tee gpurun_out/pytest_${TAG}.log
for wl in cfg2 cfg4; do
  timeout -s KILL 300 python bench.py --workload $wl --steps 20 --warmup 5 --no-e2e --no-cpu > gpurun_out/bench_${wl}_${TAG}.json 2> gpurun_out/bench_${wl}_${TAG}.err || tail -5 gpurun_out/bench_${wl}_${TAG}.err
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/bench_${wl}_${TAG}.json").read().strip().splitlines()[-1])
    print("${wl}", round(d["value"],1), round(d["ms_per_step"],4), d.get("stages_ms"), "pipeline_frac", d["roofline"].get("pipeline_frac"))
except Exception as e: print("${wl} failed", e)
PY
done
KRE='regex:fir_tmem|fir_mma|resample_generic|passthrough|cover_kernel|ranges_kernel|kept_kernel|silence_|compact_kernel|stft_mel|logmel_tc|mel_floor|logmel_init|energy_ms|remap'
timeout -s KILL 600 ncu --metrics gpu__time_duration.sum --clock-control none -k "$KRE" -c 200 --csv --log-file gpurun_out/launches_${TAG}.csv python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu > gpurun_out/ncu_launches_${TAG}.log 2>&1
timeout -s KILL 900 ncu --set full --clock-control none --import-source on -k regex:"$KFULL" -s 4 -c 3 -o gpurun_out/prof_${TAG} -f python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu > gpurun_out/ncu_full_${TAG}.log 2>&1
tail -2 gpurun_out/ncu_full_${TAG}.log
