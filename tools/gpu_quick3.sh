#!/bin/bash
# GPU tests + bench lines + a launch list of one cfg2 bench (kernel durations).  TAG=$1  WORKLOADS=$2
TAG=${1:-q}
WLS=${2:-"cfg2 cfg1 cfg4"}
bash tools/gpu_quick2.sh "$TAG" "$WLS"
timeout -s KILL 600 ncu --metrics gpu__time_duration.sum --clock-control none --kernel-name 'regex:fir_|silence|stft_mel|mel_floor|compact|passthrough|resample_generic' -c 200 --csv --log-file gpurun_out/launches_${TAG}.csv \
  python bench.py --workload cfg2 --steps 2 --warmup 1 --no-cpu --no-e2e > gpurun_out/ncu_${TAG}.log 2>&1 || tail -5 gpurun_out/ncu_${TAG}.log
python - <<PY
import csv, collections
rows = list(csv.reader(l for l in open("gpurun_out/launches_${TAG}.csv") if l.startswith('"')))
h = rows[0]; ki = h.index("Kernel Name"); vi = h.index("Metric Value"); ui = h.index("Metric Unit")
d = collections.defaultdict(list)
for r in rows[1:]:
    v = float(r[vi].replace(",", "")); v = v / 1000 if r[ui] in ("ns", "nsecond") else v
    d[r[ki][:60]].append(v)
for k, v in d.items(): print(f"{k:62s} n={len(v):3d} last={v[-1]:9.1f} us  min={min(v):9.1f}")
PY
