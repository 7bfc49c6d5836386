#!/bin/bash
# one GPU call: tests + bench of every workload + ncu launch list + full capture of the two top kernels
set -x
timeout -s KILL 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
mkdir -p gpurun_out
KRE='regex:fir_tmem|fir_umma|fir_mma|fir_fast|resample_generic|passthrough|cover_kernel|ranges_kernel|kept_kernel|compact_kernel|stft_mel|mel_floor|logmel_init|energy_ms'
TAG=${1:-r1}
for wl in cfg2 cfg1 cfg4 cfg3; do
  timeout -s KILL 600 python bench.py --workload $wl --steps 20 --warmup 5 > gpurun_out/bench_${wl}_${TAG}.json 2> gpurun_out/bench_${wl}_${TAG}.err || tail -5 gpurun_out/bench_${wl}_${TAG}.err
  cat gpurun_out/bench_${wl}_${TAG}.json
done
timeout -s KILL 300 python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu > gpurun_out/plain_${TAG}.log 2>&1 &&
timeout -s KILL 600 ncu --metrics gpu__time_duration.sum --clock-control none -k "$KRE" -c 400 --csv --log-file gpurun_out/launches_${TAG}.csv python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu > gpurun_out/ncu_launches_${TAG}.log 2>&1
timeout -s KILL 900 ncu --set full --clock-control none --import-source on -k regex:"fir_tmem_kernel|fir_mma_kernel|stft_mel_kernel|cover_kernel|compact_kernel" -s 9 -c 9 -o gpurun_out/prof_${TAG} -f python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu > gpurun_out/ncu_full_${TAG}.log 2>&1
tail -3 gpurun_out/ncu_full_${TAG}.log
ls -la gpurun_out
