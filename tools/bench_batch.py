import sys, json, numpy as np, torch
sys.path.insert(0, ".")
import openkitchen_b200 as ok
n, rays = int(sys.argv[1]), 32
env = ok.BatchEnv(ok.track_names(), n, rays=rays, reward_mode=ok.REWARD_CMAES_PROGRESS, auto_reset=1)
ids = np.arange(n, dtype=np.uint64)
pts_per = np.asarray(env.points_per_track, dtype=np.uint64)[env.track_id.numpy()]
env.reset(None, torch.as_tensor(((ids * np.uint64(2654435761)) % np.uint64(2**32) % pts_per).astype(np.int32)))
for _ in range(10): env.step_random(1)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(50): env.step_random(1)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 50
print(n, "ms/tick", ms, "rays/s", n * rays / ms * 1e3, "tiles", env.env.launch_stats().tiles, flush=True)
