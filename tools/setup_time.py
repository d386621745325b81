import sys, time
sys.path.insert(0, ".")
import numpy as np, openkitchen_b200 as ok
env = ok.Env(device=0)
t0 = time.perf_counter()
for nm in ok.track_names()[:4]:
    env.add_named_track(nm)
t1 = time.perf_counter()
env.alloc_agents(1024, ok.ray_fan(32), (np.arange(1024) * 4 // 1024).astype(np.int32))
t2 = time.perf_counter()
env.cast_rays(); env.sync()
t3 = time.perf_counter()
print(f"add tracks {t1-t0:.2f}s alloc (tables+arena) {t2-t1:.2f}s first launch {t3-t2:.2f}s")
