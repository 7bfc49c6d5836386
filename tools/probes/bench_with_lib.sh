#!/bin/bash
# bench.py lines with an alternative build of the library swapped in: bench_with_lib.sh <lib.so> <workload> [bench args...]
LIBALT=$1; WL=$2; shift 2
cp audio_processor_b200/libb2a.so /tmp/libb2a_keep.so
cp $LIBALT audio_processor_b200/libb2a.so
timeout -s KILL 300 python bench.py --workload $WL --steps 20 --warmup 5 --no-e2e --no-cpu "$@" 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$(basename $LIBALT)', '$WL', '$*', round(d['value'],1), round(d['ms_per_step'],4))"
cp /tmp/libb2a_keep.so audio_processor_b200/libb2a.so
