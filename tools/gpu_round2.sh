#!/bin/bash
# one GPU call of round 2: GPU tests, bench lines, ncu launch list + full capture of the top kernels.  TAG=$1
TAG=${1:-r02a}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.sm --format=csv,noheader
timeout -s KILL 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 | tee gpurun_out/pytest_${TAG}.log
for wl in cfg2 cfg1 cfg4 cfg3; do
  timeout -s KILL 400 python bench.py --workload $wl --steps 20 --warmup 5 > gpurun_out/bench_${wl}_${TAG}.json 2> gpurun_out/bench_${wl}_${TAG}.err || tail -5 gpurun_out/bench_${wl}_${TAG}.err
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/bench_${wl}_${TAG}.json").read().strip().splitlines()[-1])
    print("${wl}", d["value"], d["ms_per_step"], d.get("stages_ms"), d.get("roofline"), d.get("e2e"))
except Exception as e: print("${wl} failed", e)
PY
done
KRE='regex:fir_tmem|fir_mma|resample_generic|passthrough|cover_kernel|ranges_kernel|kept_kernel|silence_|compact_kernel|stft_mel|logmel_tc|mel_floor|logmel_init|energy_ms|remap'
timeout -s KILL 300 python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu > gpurun_out/plain_${TAG}.log 2>&1 &&
timeout -s KILL 600 ncu --metrics gpu__time_duration.sum --clock-control none -k "$KRE" -c 400 --csv --log-file gpurun_out/launches_${TAG}.csv python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu > gpurun_out/ncu_launches_${TAG}.log 2>&1
timeout -s KILL 900 ncu --set full --clock-control none --import-source on -k regex:"logmel_tc_kernel|fir_tmem_kernel|cover_kernel|silence_" -s 6 -c 6 -o gpurun_out/prof_${TAG} -f python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu > gpurun_out/ncu_full_${TAG}.log 2>&1
tail -3 gpurun_out/ncu_full_${TAG}.log
ls -la gpurun_out | tail -20
