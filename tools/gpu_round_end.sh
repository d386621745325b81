#!/bin/bash
# what the driver runs at round end, plus the evidence files: tools/gpu_round_end.sh <tag>
cd "$(dirname "$0")/.."
tag=${1:-r6}
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/pytest_gpu_$tag.log 2>&1
grep -E "passed|failed|error" gpurun_out/pytest_gpu_$tag.log | tail -2
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
tools/checked_build.sh
timeout 900 python bench.py > gpurun_out/bench_$tag.json 2> gpurun_out/bench_$tag.err; echo "bench exit $? lines $(wc -l < gpurun_out/bench_$tag.json)"
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref_$tag.json 2> gpurun_out/bench_ref_$tag.err; echo "ref exit $?"
timeout 900 python tools/bench_configs.py > gpurun_out/cfg_$tag.jsonl 2> gpurun_out/cfg_$tag.err; echo "cfg exit $?"
