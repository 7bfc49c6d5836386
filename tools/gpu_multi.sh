#!/bin/bash
# multi-GPU bench lines: N=$1 ranks, workloads $2 (default: the N>1 default cfg5, then cfg4 strong scaling)
N=${1:-8}
WLS=${2:-"cfg5 cfg4"}
mkdir -p gpurun_out
for wl in $WLS; do
  timeout -s KILL 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 \
    bench.py --gpus $N --workload $wl --steps 10 --warmup 3 > gpurun_out/bench_${wl}_n${N}.json 2> gpurun_out/bench_${wl}_n${N}.err || tail -5 gpurun_out/bench_${wl}_n${N}.err
  tail -1 gpurun_out/bench_${wl}_n${N}.json | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('$wl', 'n_gpus', d['n_gpus'], round(d['value'],1), d['unit'], round(d['ms_per_step'],3), 'ms', 'scaling', d['scaling'], 'e2e', d['e2e'].get('value'), d['e2e'].get('frac_of_h2d_ceiling'), 'clocks', d.get('clocks'))"
done
