cd /root/repo
for g in 2 4 8 16 32 296; do echo -n "store probe, $g CTAs: "; OK_PROBE_GRID=$g python -c "
import openkitchen_b200 as ok
print(round(ok.pcie_probe(0, 8716288, 50, 'store'),1), 'GB/s')"; done
run() { echo -n "$* : ";  env "$@" OK_BENCH_CONFIG5=0 python bench.py --steps 100 --warmup 20 --no-cpu 2>/dev/null | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(round(d['ms_per_step'],4), round(d['e2e']['ms_per_step'],4))"; }
run OK_E2E_FLUSHERS=0
run OK_E2E_FLUSHERS=4 OK_E2E_TILES=4
run OK_E2E_FLUSHERS=8 OK_E2E_TILES=4
run OK_E2E_FLUSHERS=12 OK_E2E_TILES=4
run OK_E2E_FLUSHERS=8 OK_E2E_TILES=2
run OK_E2E_FLUSHERS=8 OK_E2E_TILES=6
run OK_E2E_FLUSHERS=16 OK_E2E_TILES=6
(timeout 300 python -m pytest tests/test_gpu_parity_gaps.py -q -x -k host_step) 2>&1 | tail -3
