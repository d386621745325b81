#!/usr/bin/env python
"""Where one tick of the PPO rollout (config 4: 4,096 envs x 5 rays, actor 5-128-3) goes: per-tile phase times of the
fused actor + step kernel (ok_debug_trace) next to the per-tick time of the captured rollout.  usage: python tools/c4_timeline.py"""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import openkitchen_b200 as ok  # noqa: E402
from openkitchen_b200.rollout import FusedActorRollout  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
torch.manual_seed(0)
l1, l2 = torch.nn.Linear(5, 128).cuda(), torch.nn.Linear(128, 3).cuda()
table = torch.tensor([[60.0, 0.0], [60.0, -3.0], [60.0, 3.0]], device="cuda")
for one_launch in (True, False):
    env = ok.BatchEnv(["Monza"], n, rays=[-70, -30, 0, 30, 70], reward_mode=ok.REWARD_CONSTANT, auto_reset=1)
    fr = FusedActorRollout(env, l1, l2, table, 64, sample=True, one_launch=one_launch)
    env.reset_random()
    env.cast_rays()
    fr.run()
    torch.cuda.synchronize()
    res = []
    for rep in range(4):
        env.env.debug_trace(4)
        fr._tick(rep)
        torch.cuda.synchronize()
        tr = env.env.debug_trace(0)
        used = tr[:, :, 1] > 0
        t0 = tr[:, :, 1][used].min()
        t = (tr[:, :, 1:].astype(np.int64) - int(t0)) / 1e3
        res.append({"start_max": float(t[:, :, 0][used].max()), "phase0+1_end": float(t[:, :, 1][used].mean()), "passA_end": float(t[:, :, 2][used].mean()),
                    "rays_end": float(t[:, :, 3][used].mean()), "tile_end_mean": float(t[:, :, 4][used].mean()), "tile_end_max": float(t[:, :, 4][used].max())})
    print(json.dumps({"one_launch": one_launch, "trace_us": {k: round(float(np.median([r[k] for r in res])), 2) for k in res[0]}}), flush=True)
