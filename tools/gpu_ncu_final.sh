#!/bin/bash
# the two ncu passes behind profiles/r2_step_kernel_* and profiles/traffic.json: tools/gpu_ncu_final.sh <tag>
cd "$(dirname "$0")/.."
tag=${1:-r7}
mkdir -p gpurun_out
timeout 300 python bench.py --steps 40 --warmup 10 --no-cpu > gpurun_out/bench_short_$tag.json 2> gpurun_out/bench_short_$tag.err; echo "plain run exit $?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 30 -c 200 --csv --log-file gpurun_out/launches_$tag.csv \
  python bench.py --steps 40 --warmup 10 --no-cpu > gpurun_out/launches_$tag.log 2>&1; echo "launch list exit $?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:step_kernel -s 90 -c 1 -o gpurun_out/prof_$tag -f \
  python bench.py --steps 20 --warmup 30 --no-cpu > gpurun_out/prof_$tag.log 2>&1; echo "ncu exit $?"
