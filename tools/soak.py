#!/usr/bin/env python
"""Soak test: many ticks of the bench workload at several population sizes, under a watchdog.
usage: timeout 300 python tools/soak.py [steps]"""
import sys, time
sys.path.insert(0, '.')
import bench, openkitchen_b200 as ok
steps = int(sys.argv[1]) if len(sys.argv) > 1 else 300
for n in (2300, 23000, 65536):
    env = ok.Env(device=0, movement_mode=0, reward_mode=2, auto_reset=1)
    bench.build_workload(ok, env, n)
    t0 = time.time()
    for s in range(0, steps, 20):
        env.launch_steps_random(s, 20)
        env.sync()
    print(f"n={n} {steps} ticks ok in {time.time()-t0:.2f}s", flush=True)
    env.close()
