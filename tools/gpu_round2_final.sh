#!/bin/bash
# round-2 evidence run on ONE GPU: GPU tests, bench lines of every workload (with the CPU baseline), ncu launch list and full
# captures of the top kernels.  Outputs under gpurun_out/ (copied to profiles/ by hand).  TAG=$1
TAG=${1:-r02}
mkdir -p gpurun_out
timeout -s KILL 900 python -m pytest tests -m gpu -q 2>&1 | tail -6 | tee gpurun_out/pytest_${TAG}.log
for wl in cfg2 cfg1 cfg3 cfg4 cfg5; do
  timeout -s KILL 900 python bench.py --workload $wl --steps 50 --warmup 5 > gpurun_out/${TAG}_bench_${wl}.json 2> gpurun_out/${TAG}_bench_${wl}.err || tail -5 gpurun_out/${TAG}_bench_${wl}.err
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/${TAG}_bench_${wl}.json").read().strip().splitlines()[-1])
    print("${wl}", round(d["value"],1), round(d["ms_per_step"],4), d.get("stages_ms"), round(d["roofline"]["frac"],3), round(d["roofline"]["pipeline_frac"],3), (d.get("e2e") or {}).get("value"), (d.get("cpu_baseline") or {}).get("value"))
except Exception as e: print("${wl} failed", e)
PY
done
timeout -s KILL 900 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${TAG}_bench_reference.json 2>/dev/null; tail -c 600 gpurun_out/${TAG}_bench_reference.json
KRE='regex:fir_tmem|fir_mma|resample_generic|passthrough|silence_|compact_kernel|stft_mel|logmel_tc|mel_floor|logmel_init|energy_ms|remap|kept_offsets|mel_windows'
timeout -s KILL 600 ncu --metrics gpu__time_duration.sum --clock-control none -k "$KRE" -c 300 --csv --log-file gpurun_out/${TAG}_ncu_launches.csv python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu > gpurun_out/${TAG}_ncu_launches.log 2>&1
timeout -s KILL 900 ncu --set full --clock-control none --import-source on -k regex:"stft_mel_kernel|fir_tmem_kernel|silence_kernel|mel_floor_kernel" -s 8 -c 8 -o gpurun_out/${TAG}_prof -f python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu > gpurun_out/${TAG}_ncu_full.log 2>&1
tail -2 gpurun_out/${TAG}_ncu_full.log
timeout -s KILL 600 ncu --set full --clock-control none -k regex:"stft_mel_kernel" -s 1 -c 1 -o gpurun_out/${TAG}_prof_cfg3 -f python bench.py --workload cfg3 --steps 1 --warmup 3 --no-e2e --no-cpu > gpurun_out/${TAG}_ncu_full_cfg3.log 2>&1
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.sm,power.draw --format=csv,noheader
