#!/bin/bash
# ncu captures of the step kernel: usage tools/gpu_prof.sh <tag> <bench args...>
cd "$(dirname "$0")/.."
tag=$1; shift
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"step_kernel|beam_cast_kernel" -s 90 -c 1 -o gpurun_out/prof_$tag -f \
  python bench.py --steps 20 --warmup 30 --no-cpu "$@" > gpurun_out/prof_$tag.log 2>&1
echo "ncu $tag exit $?"
