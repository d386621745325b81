#!/bin/bash
# 8-GPU evidence: the bench line at N = 8 (NCCL fitness exchange in the timed arm, shard parity, config-5 leg) and the
# end-to-end path under its transfer switches; usage: tools/gpu8.sh <tag> [n_gpus]
cd "$(dirname "$0")/.."
tag=${1:-r2}; n=${2:-8}
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus $n \
  > gpurun_out/bench${n}_$tag.json 2> gpurun_out/bench${n}_$tag.err; echo "bench N=$n exit $?"
if [ -n "$OK_GPU8_MAIN_ONLY" ]; then set --; else set -- "OK_HOST_OBS=copy" "OK_E2E_TILES=2" "OK_E2E_TILES=8"; fi
for v in "$@"; do
  env $v OK_BENCH_CONFIG5=0 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29542 bench.py --gpus $n --steps 100 --warmup 20 --no-cpu 2>/dev/null | tail -1 > gpurun_out/bench${n}_${tag}_$(echo $v | tr '=' '_').json
  echo "$v: $(python -c "import json; d=json.loads(open('gpurun_out/bench${n}_${tag}_$(echo $v | tr '=' '_').json').read()); print(d['ms_per_step'], d['e2e']['ms_per_step'], d['e2e']['host_link'])")"
done
python - <<PY
import json
d = json.loads(open("gpurun_out/bench${n}_$tag.json").read().strip().splitlines()[-1])
print("value", d["value"], "ms", d["ms_per_step"], "e2e", d["e2e"]["value"], d["e2e"]["ms_per_step"], d["e2e"]["host_link"])
print("generation", d.get("generation")); print("shard_parity", d.get("shard_parity")); print("config5", d.get("config5")); print("numa", d["impl_config"]["numa"]); print("cpu", d.get("cpu_baseline"))
PY
nvidia-smi topo -m 2>/dev/null | head -14
