"""experiment (variant build -DOK_PRE_PROBE): where thread 0 of every CTA spends its phase 1.
usage: OK_B200_LIB=openkitchen_b200/lib/variants/lib_probe.so python tools/probe_phase1.py"""
import sys
import numpy as np
import torch
sys.path.insert(0, ".")
import bench  # noqa: E402
import openkitchen_b200 as ok  # noqa: E402

env = ok.Env(device=0, reward_mode=ok.REWARD_CMAES_PROGRESS, auto_reset=1)
bench.build_workload(ok, env, 65536)
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
env.launch_steps_random(0, 300, bench.SEED)
env.sync()
for rep in range(3):
    env.debug_trace(1)
    flush.zero_()
    torch.cuda.synchronize()
    env.launch_steps_random(300 + rep, 1, bench.SEED, torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    tr = env.debug_trace(0)[:, 0, :].astype(np.int64)
    t0 = tr[:, 1].min()
    names = ["tile start", "phase 1 done (CTA)", "state loaded", "kinematics + sincos done", "table row loaded"]
    cols = [1, 2, 3, 4, 5]
    print({n: round(float((tr[:, c] - t0).mean()) / 1e3, 2) for n, c in zip(names, cols)})
