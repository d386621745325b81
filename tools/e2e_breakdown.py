#!/usr/bin/env python
"""Where ok_step_host's tick goes (C3 shape, one GPU): the host-synchronous loop of bench.py's e2e leg with the outputs
switched on one at a time and the end-to-end tiling (OK_E2E_TILES tiles per CTA) varied, for float and q16 observations.
usage: python tools/e2e_breakdown.py [steps]"""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import bench  # noqa: E402
import openkitchen_b200 as ok  # noqa: E402

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 200
n, R = bench.N_AGENTS, bench.N_RAYS
rng = np.random.default_rng(1)
thr = ok.pinned_array((steps, n), np.float32)
steer = ok.pinned_array((steps, n), np.float32)
thr[:] = rng.random((steps, n), dtype=np.float32) * 100.0
steer[:] = rng.random((steps, n), dtype=np.float32) * 10.0 - 5.0
obs = ok.pinned_array((n, R), np.float32)
obs_q = ok.pinned_array((n, R), np.uint16)
rew = ok.pinned_array((n,), np.float32)
done = ok.pinned_array((n,), np.uint8)


def run(env, a, b, o, r, d):
    for i in range(5):
        env.step_host(thr[i] if a else None, steer[i] if b else None, o, r, d)
    t = time.perf_counter()
    for i in range(steps):
        env.step_host(thr[i] if a else None, steer[i] if b else None, o, r, d)
    return 1e3 * (time.perf_counter() - t) / steps


for tiles in [int(x) for x in os.environ.get("TILES", "1,2,3,4,6").split(",")]:
    os.environ["OK_E2E_TILES"] = str(tiles)
    env = ok.Env(device=0, movement_mode=ok.MOVE_VELOCITY, reward_mode=ok.REWARD_CMAES_PROGRESS, auto_reset=1)
    bench.build_workload(ok, env, n, id_base=0, period=n)
    env.launch_steps_random(0, 300, bench.SEED)
    env.sync()
    out = {"tiles_per_cta": tiles}
    if tiles == 1:
        out["launch+sync only (stored actions)"] = run(env, False, False, None, None, None)
        out["+ mapped actions"] = run(env, True, True, None, None, None)
        out["+ reward/done"] = run(env, True, True, None, rew, done)
    out["f32 obs + reward/done"] = run(env, True, True, obs, rew, done)
    out["q16 obs + reward/done"] = run(env, True, True, obs_q, rew, done)
    out["f32 obs only"] = run(env, True, True, obs, None, None)
    print(json.dumps({k: (round(v, 4) if isinstance(v, float) else v) for k, v in out.items()}), flush=True)
    del env
