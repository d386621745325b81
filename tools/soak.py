"""long full-size comparison of the three narrow phases: beam lists vs brute force (and grid) on every buffer
usage: python tools/soak.py [ticks] [check_every] [rays] [movement_mode 0|1]"""
import sys
import time

import numpy as np

sys.path.insert(0, ".")
import bench  # noqa: E402
import openkitchen_b200 as ok  # noqa: E402

ticks = int(sys.argv[1]) if len(sys.argv) > 1 else 300
every = int(sys.argv[2]) if len(sys.argv) > 2 else 50
bench.N_RAYS = int(sys.argv[3]) if len(sys.argv) > 3 else 32
movement = int(sys.argv[4]) if len(sys.argv) > 4 else ok.MOVE_VELOCITY
envs = {}
for name, mode in (("beam", ok.RAYCAST_BEAM), ("brute", ok.RAYCAST_BRUTE), ("grid", ok.RAYCAST_GRID)):
    e = ok.Env(device=0, movement_mode=movement, reward_mode=ok.REWARD_CMAES_PROGRESS, auto_reset=1, raycast_mode=mode)
    bench.build_workload(ok, e, 65536)
    envs[name] = e
t0 = time.time()
bad = 0
for start in range(0, ticks, every):
    for e in envs.values():
        e.launch_steps_random(start, every)
    for e in envs.values():
        e.sync()
    ref = envs["brute"]
    for name in ("beam", "grid"):
        for buf in ok.BUFFERS:
            a, b = envs[name].read(buf), ref.read(buf)
            if not np.array_equal(a.view(np.uint8), b.view(np.uint8)):
                bad += 1
                print(f"MISMATCH after {start + every} ticks: {name} vs brute, buffer {buf}, {(a != b).sum()} elements", flush=True)
    print(f"{start + every} ticks ok so far: {bad == 0}  ({time.time() - t0:.0f} s), crashed {envs['beam'].read('crashed').mean():.4f}", flush=True)
print("SOAK", "PASSED" if bad == 0 else "FAILED", ticks, "ticks x", 65536 * bench.N_RAYS, "rays, movement mode", movement)
