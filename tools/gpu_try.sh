#!/bin/bash
# one gpurun call: GPU tests, then bench variants; everything logged under gpurun_out/
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nproc > gpurun_out/nproc.txt
( time timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log
for cfg in "beam 8 64" "grid 8 64" "beam 8 128" "beam 4 64" "beam 4 128"; do
  set -- $cfg
  ( time timeout 400 python bench.py --steps 300 --warmup 30 --no-cpu --raycast $1 --beam-cell $2 --beam-bins $3 ) > gpurun_out/bench_$1_$2_$3.json 2> gpurun_out/bench_$1_$2_$3.err
  echo "$cfg: $(python -c "import json,sys; d=json.loads(open('gpurun_out/bench_$1_$2_$3.json').read().strip().splitlines()[-1]); print(d['ms_per_step'], d['ray_casts_per_sec'], d['e2e']['ms_per_step'], d['config'].get('setup_s'))" 2>&1 | tail -1)"
done
