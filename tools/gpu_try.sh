#!/bin/bash
# one gpurun call: GPU tests, then bench variants; everything logged under gpurun_out/
# usage: tools/gpu_try.sh [tag] ["mode cell bins" ...]
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
tag=${1:-try}; shift
( time timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/pytest_gpu_$tag.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu_$tag.log
grep -E "passed|failed|error" gpurun_out/pytest_gpu_$tag.log | tail -3
if [ $# -eq 0 ]; then set -- "beam 8 64" "beam 4 128"; fi
for cfg in "$@"; do
  set -- $cfg
  f=gpurun_out/bench_${tag}_$1_$2_$3
  ( time timeout 400 python bench.py --steps 300 --warmup 30 --no-cpu --raycast $1 --beam-cell $2 --beam-bins $3 ) > $f.json 2> $f.err
  echo "$cfg: $(python -c "import json,sys; d=json.loads(open('$f.json').read().strip().splitlines()[-1]); print(d['ms_per_step'], d['ray_casts_per_sec'], d['e2e']['ms_per_step'], d['config'].get('setup_s'))" 2>&1 | tail -1)"
done
