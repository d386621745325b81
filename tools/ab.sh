#!/bin/bash
# A/B of two builds of the library on one box: tools/ab.sh <variant .so> [rounds]   (the default build is the other arm)
cd "$(dirname "$0")/.."
var=$1; n=${2:-2}
mkdir -p gpurun_out
for i in $(seq 1 $n); do
  for arm in new old; do
    if [ $arm = old ]; then export OK_B200_LIB=$PWD/$var; else unset OK_B200_LIB; fi
    timeout 300 python bench.py --no-cpu 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.readline()); print('$arm', $i, 'ms_per_tick', round(d['ms_per_step'],5), 'rays/s', '%.4g'%d['ray_casts_per_sec'], 'e2e_ms', round(d['e2e']['ms_per_step'],4))"
  done
done
