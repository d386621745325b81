#!/bin/bash
# round-end evidence: tests, full bench line, reference arm, config table, launch list and one full ncu capture of the step kernel
cd "$(dirname "$0")/.."
tag=${1:-r2}
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -q ) > gpurun_out/pytest_gpu_$tag.log 2>&1
grep -E "passed|failed|error" gpurun_out/pytest_gpu_$tag.log | tail -2
tools/checked_build.sh
timeout 900 python bench.py > gpurun_out/bench_$tag.json 2> gpurun_out/bench_$tag.err; echo "bench exit $?"
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref_$tag.json 2> gpurun_out/bench_ref_$tag.err; echo "ref exit $?"
timeout 900 python tools/bench_configs.py > gpurun_out/cfg_$tag.jsonl 2> gpurun_out/cfg_$tag.err; echo "cfg exit $?"
timeout 300 python tools/trace_timeline.py > gpurun_out/timeline_$tag.json 2> gpurun_out/timeline_$tag.err; echo "timeline exit $?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 30 -c 200 --csv --log-file gpurun_out/launches_$tag.csv \
  python bench.py --steps 40 --warmup 10 --no-cpu > gpurun_out/launches_$tag.log 2>&1; echo "launch list exit $?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:step_kernel -s 90 -c 1 -o gpurun_out/prof_$tag -f \
  python bench.py --steps 20 --warmup 30 --no-cpu > gpurun_out/prof_$tag.log 2>&1; echo "ncu exit $?"
