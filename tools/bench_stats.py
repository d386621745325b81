"""ms/tick + first-pass statistics of the beam kernel at the bench workload for a table resolution
usage: python tools/bench_stats.py CELL BINS [agents]"""
import json
import sys
import time

import torch

sys.path.insert(0, ".")
import bench  # noqa: E402
import openkitchen_b200 as ok  # noqa: E402

cell, bins = float(sys.argv[1]), int(sys.argv[2])
n = int(sys.argv[3]) if len(sys.argv) > 3 else 65536
t0 = time.time()
env = ok.Env(device=0, reward_mode=ok.REWARD_CMAES_PROGRESS, auto_reset=1, beam_cell=cell, beam_bins=bins)
bench.build_workload(ok, env, n)
env.cast_rays()
env.sync()
setup = time.time() - t0
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
st = torch.cuda.current_stream()
env.launch_steps_random(0, 100, bench.SEED, st.cuda_stream)
ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(300)]
for i, (a, b) in enumerate(ev):
    flush.zero_()
    a.record(st)
    env.launch_steps_random(100 + i, 1, bench.SEED, st.cuda_stream)
    b.record(st)
torch.cuda.synchronize()
ms = sum(a.elapsed_time(b) for a, b in ev) / len(ev)
try:
    env.debug_stats(True)
    env.launch_steps_random(400, 50, bench.SEED, st.cuda_stream)
    env.sync()
    rays, queued, walked = env.debug_stats(False)
except AttributeError:  # an older library
    rays, queued, walked = 1, 0, 0
tb = sum(env.beam_table_bytes(t) for t in range(env.num_tracks()))
print(json.dumps({"cell": cell, "bins": bins, "agents": n, "ms_per_tick": ms, "ray_casts_per_s": n * 32 / (ms * 1e-3), "setup_s": setup,
                  "table_GB": tb / 1e9, "queued_frac": queued / max(rays, 1), "walk_frac": walked / max(rays, 1)}), flush=True)
