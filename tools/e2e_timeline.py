#!/usr/bin/env python
"""Per-tile timeline of ONE ok_step_host tick (ok_debug_trace): when each of a CTA's tiles starts, finishes its rays and
leaves phase 4, with float / q16 / no host observations.  usage: [TILES=4] python tools/e2e_timeline.py"""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import bench  # noqa: E402
import openkitchen_b200 as ok  # noqa: E402

n, R = bench.N_AGENTS, bench.N_RAYS
tiles_per_cta = int(os.environ.get("TILES", "4"))
os.environ["OK_E2E_TILES"] = str(tiles_per_cta)
env = ok.Env(device=0, reward_mode=ok.REWARD_CMAES_PROGRESS, auto_reset=1)
bench.build_workload(ok, env, n)
env.launch_steps_random(0, 300, bench.SEED)
env.sync()
rng = np.random.default_rng(1)
thr, steer = ok.pinned_array((n,), np.float32), ok.pinned_array((n,), np.float32)
thr[:] = rng.random(n, dtype=np.float32) * 100.0
steer[:] = rng.random(n, dtype=np.float32) * 10.0 - 5.0
obs, obs_q = ok.pinned_array((n, R), np.float32), ok.pinned_array((n, R), np.uint16)
rew, done = ok.pinned_array((n,), np.float32), ok.pinned_array((n,), np.uint8)
sp = torch.cuda.current_stream().cuda_stream
for name, o in (("none", None), ("f32", obs), ("q16", obs_q)):
    for rep in range(3):
        env.step_host(thr, steer, o, rew, done, sp)
    env.debug_trace(16)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    env.step_host(thr, steer, o, rew, done, sp)
    e1.record()
    torch.cuda.synchronize()
    tr = env.debug_trace(0)
    used = tr[:, :, 1] > 0
    t0 = tr[:, :, 1][used].min()
    t = (tr[:, :, 1:].astype(np.int64) - int(t0)) / 1e3
    out = {"obs": name, "tiles_per_cta": tiles_per_cta, "event_ms": round(e0.elapsed_time(e1), 4), "ctas": int(tr.shape[0]), "tiles": int(used.sum())}
    for k in range(min(tr.shape[1], 8)):
        u = used[:, k]
        if not u.any():
            break
        out[f"tile{k}"] = {"n": int(u.sum()), "start": round(float(t[u, k, 0].mean()), 1), "p1": round(float(t[u, k, 1].mean()), 1),
                           "rays_end": round(float(t[u, k, 3].mean()), 1), "end": round(float(t[u, k, 4].mean()), 1),
                           "end_max": round(float(t[u, k, 4].max()), 1)}
    out["last_end_us"] = round(float(np.where(used, t[:, :, 4], 0).max()), 1)
    print(json.dumps(out), flush=True)
