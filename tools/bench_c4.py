#!/usr/bin/env python
"""BASELINE configs[3] (PPO rollout collection: 4,096 envs x 256 steps, actor 5-128-3) through FusedActorRollout.  usage: python tools/bench_c4.py [envs] [steps]"""
import json
import os
import sys
import time

import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import openkitchen_b200 as ok  # noqa: E402
from openkitchen_b200.rollout import FusedActorRollout, discounted_returns_fused  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 256
torch.manual_seed(0)
l1, l2 = torch.nn.Linear(5, 128).cuda(), torch.nn.Linear(128, 3).cuda()
table = torch.tensor([[60.0, 0.0], [60.0, -3.0], [60.0, 3.0]], device="cuda")
for rep in range(2):
    env = ok.BatchEnv(["Monza"], n, rays=[-70, -30, 0, 30, 70], reward_mode=ok.REWARD_CONSTANT, auto_reset=1)
    fr = FusedActorRollout(env, l1, l2, table, steps, sample=True)
    env.reset_random()
    env.cast_rays()
    fr.capture()
    fr.run()
    torch.cuda.synchronize()
    reps = 20
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record()
    for i in range(reps):
        fr.run()
        ret = discounted_returns_fused(env, fr.rewards, 0.99, fr.dones)
    e1.record()
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / reps
    print(json.dumps({"envs": n, "steps": steps, "ms_per_tick": round(1e3 * dt / steps, 5),
                      "device_ms_per_tick": round(e0.elapsed_time(e1) / reps / steps, 5), "env_steps_per_sec": n * steps / dt,
                      "done_fraction": float(fr.dones.float().mean())}), flush=True)
    del fr, env
