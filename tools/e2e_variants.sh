#!/bin/bash
# end-to-end (ok_step_host) ms per tick under the transfer switches; usage: tools/e2e_variants.sh <n_gpus>
cd "$(dirname "$0")/.."
n=${1:-1}
run() {
  if [ "$n" = 1 ]; then env "$@" OK_BENCH_CONFIG5=0 python bench.py --steps 100 --warmup 20 --no-cpu 2>/dev/null
  else env "$@" OK_BENCH_CONFIG5=0 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $n --steps 100 --warmup 20 --no-cpu 2>/dev/null
  fi | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(round(d['ms_per_step'],4), round(d['e2e']['ms_per_step'],4), d['e2e']['host_link']['d2h_copy_GBps_slowest_rank'], d['e2e']['host_link']['mapped_store_GBps_slowest_rank'])"
}
echo "default:";            run A=1
echo "OK_HOST_SMALL=copy:"; run OK_HOST_SMALL=copy
echo "OK_HOST_ACT=copy:";   run OK_HOST_ACT=copy
echo "OK_HOST_OBS=copy:";   run OK_HOST_OBS=copy
