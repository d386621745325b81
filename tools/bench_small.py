"""ms per tick of the two beam-kernel shapes over population size (back-to-back ticks, no L2 flush: the latency view)
usage: python tools/bench_small.py [rays]"""
import json
import os
import sys

import torch

sys.path.insert(0, ".")
import bench  # noqa: E402
import openkitchen_b200 as ok  # noqa: E402

rays = int(sys.argv[1]) if len(sys.argv) > 1 else 32
bench.N_RAYS = rays
for n in (1, 64, 1024, 4096, 8192, 16384, 32768, 65536):
    row = {"agents": n, "rays": rays}
    for kernel in ("staged", "unstaged"):
        os.environ["OK_BEAM_KERNEL"] = kernel
        env = ok.Env(device=0, reward_mode=ok.REWARD_CMAES_PROGRESS, auto_reset=1)
        bench.build_workload(ok, env, n)
        st = torch.cuda.current_stream()
        env.launch_steps_random(0, 50, bench.SEED, st.cuda_stream)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(st)
        env.launch_steps_random(50, 300, bench.SEED, st.cuda_stream)
        e1.record(st)
        torch.cuda.synchronize()
        row[kernel + "_us"] = 1e3 * e0.elapsed_time(e1) / 300
        row[kernel + "_tiles"] = env.launch_stats().tiles
        env.close()
    print(json.dumps(row), flush=True)
