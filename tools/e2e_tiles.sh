#!/bin/bash
# ok_step_host ms per tick over the number of tiles per CTA of its schedule; usage: tools/e2e_tiles.sh
cd "$(dirname "$0")/.."
for t in 1 2 3 4 6 8; do
  echo -n "OK_E2E_TILES=$t: "
  OK_E2E_TILES=$t OK_BENCH_CONFIG5=0 python bench.py --steps 100 --warmup 20 --no-cpu 2>/dev/null | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(round(d['ms_per_step'],4), round(d['e2e']['ms_per_step'],4))"
done
