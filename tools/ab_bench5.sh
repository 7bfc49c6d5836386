#!/bin/bash
# A/B of two libraries on the many-clip workloads (short runs): working-tree libb2a.so against $1
ALT=${1:-tools/probes/_bin/libb2a_head.so}; shift
WLS=${@:-cfg5}
cp audio_processor_b200/libb2a.so /tmp/libb2a_new.so
line() { python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$1', d['config']['workload'].split(':')[0], round(d['value'],1), round(d['ms_per_step'],4))"; }
for rep in 1 2; do
  for wl in $WLS; do
    cp $ALT audio_processor_b200/libb2a.so; timeout -s KILL 300 python bench.py --workload $wl --steps 6 --warmup 2 --no-e2e --no-cpu 2>/dev/null | line alt
    cp /tmp/libb2a_new.so audio_processor_b200/libb2a.so; timeout -s KILL 300 python bench.py --workload $wl --steps 6 --warmup 2 --no-e2e --no-cpu 2>/dev/null | line new
  done
done
