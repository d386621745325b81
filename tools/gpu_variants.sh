#!/bin/bash
# bench several builds of the library (OK_B200_LIB) on the same box: usage tools/gpu_variants.sh <bench args> -- lib1 lib2 ...
cd "$(dirname "$0")/.."
args=(); while [ "$1" != "--" ]; do args+=("$1"); shift; done; shift
for lib in "$@"; do
  r=$(OK_B200_LIB=$PWD/openkitchen_b200/lib/$lib python bench.py --steps 300 --warmup 30 --no-cpu "${args[@]}" 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['ms_per_step'], d['e2e']['ms_per_step'])")
  echo "$lib: $r"
done
