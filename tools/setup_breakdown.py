"""where the set-up time of the bench workload goes (OK_BEAM_VERBOSE=1 adds the table builder's per-track lines)
usage: OK_BEAM_VERBOSE=1 python tools/setup_breakdown.py"""
import sys
import time

import numpy as np

sys.path.insert(0, ".")
t00 = time.perf_counter()
import torch  # noqa: E402,F401

torch.zeros(1, device="cuda")
t0 = time.perf_counter()
import openkitchen_b200 as ok  # noqa: E402

env = ok.Env(device=0, reward_mode=ok.REWARD_CMAES_PROGRESS, auto_reset=1)
t1 = time.perf_counter()
names = ok.track_names()
cols = [ok.track_columns(nm) for nm in names]
t2 = time.perf_counter()
for c in cols:
    env.add_track(c)
t3 = time.perf_counter()
n = 65536
tid = (np.arange(n) * len(names) // n).astype(np.int32)
env.alloc_agents(n, ok.ray_fan(32), tid)
t4 = time.perf_counter()
env.cast_rays()
env.sync()
t5 = time.perf_counter()
env.launch_steps_random(0, 1)
env.sync()
t6 = time.perf_counter()
print(f"torch+context {t0 - t00:.2f}s | load lib + create env {t1 - t0:.3f}s | read 23 track files {t2 - t1:.3f}s | add_track x23 (host geometry, grid) {t3 - t2:.3f}s | "
      f"alloc_agents (beam tables + arena + agents + reset) {t4 - t3:.3f}s | first cast_rays {t5 - t4:.3f}s | first step {t6 - t5:.4f}s | total after context {t6 - t0:.2f}s")
