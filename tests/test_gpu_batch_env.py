"""GPU: BatchEnv -- zero-copy DLPack views, device-side actions from a torch policy, resets without host syncs."""
import numpy as np
import pytest

import openkitchen_b200 as ok
from oracle.api import Oracle

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")


def test_views_are_zero_copy_and_match_oracle():
    env = ok.BatchEnv(["Monza", "Spa"], 128, rays=5, movement_mode=ok.MOVE_VELOCITY, reward_mode=ok.REWARD_CONSTANT)
    ptr, shape, _ = env.env.buffer_info("obs")
    assert env.obs.data_ptr() == ptr and tuple(env.obs.shape) == shape and env.obs.device.type == "cuda"
    ora = Oracle("port", movement_mode=0, reward_mode=3)
    for nm in ("Monza", "Spa"):
        ora.add_track(ok.track_columns(nm))
    ora.alloc_agents(128, ok.ray_fan(5), env.track_id.numpy())
    # a tiny deterministic "policy" on the GPU: steer away from the nearer side, throttle from the front ray
    env.cast_rays()
    ora.cast_rays()
    for _ in range(50):
        o = env.obs
        thr = 20.0 + 60.0 * o[:, 2]
        steer = 4.0 * torch.sign(o[:, 4] - o[:, 0])
        obs, rew, done = env.step(thr, steer)
        ora.step(thr.cpu().numpy(), steer.cpu().numpy())
    torch.cuda.synchronize()
    assert np.array_equal(env.obs.cpu().numpy().view(np.uint32), ora.buffer("obs").view(np.uint32))
    assert np.array_equal(env.done.cpu().numpy(), ora.buffer("done"))
    assert np.array_equal(env.pose[:, 0].cpu().numpy().view(np.uint32), ora.buffer("pos_x").view(np.uint32))
    assert float(env.reward.sum()) == 128.0
    # writing through the view is writing the env's state
    env.act_throttle.fill_(7.0)
    assert float(env.env.read("act_throttle")[5]) == 7.0


def test_reset_mask_and_random_resets():
    env = ok.BatchEnv(["Silverstone"], 256, rays=32, reward_mode=ok.REWARD_CMAES_PROGRESS)
    env.step_random(k=120)
    crashed = env.crashed.bool().clone()
    assert crashed.any()
    x_before = env.pos_x.clone()
    env.reset(crashed)  # RaceTrack::kStartingIdx
    torch.cuda.synchronize()
    assert not env.crashed.bool()[crashed].any()
    assert torch.equal(env.pos_x[~crashed], x_before[~crashed])
    track_x = torch.as_tensor(env.env.track_array(0, "x"))
    assert torch.all(env.pos_x[crashed].cpu() == track_x[3])
    g = torch.Generator(device="cuda").manual_seed(1)
    env.reset_random(generator=g, randomize_lane=True, randomize_heading=True)
    env.cast_rays()
    torch.cuda.synchronize()
    assert env.pos_x.unique().numel() > 100 and not env.crashed.any()
    env.close()


def test_views_alive_at_interpreter_exit_do_not_crash():
    """a BatchEnv left in a module global is torn down after Python: the DLPack deleters must not call back"""
    import os
    import subprocess
    import sys

    code = ("import torch, openkitchen_b200 as ok\n"
            "env = ok.BatchEnv(['Monza'], 64, rays=5)\n"
            "env.step_random(2)\n"
            "keep = [env.obs, env.reward, env['pos_x']]\n"
            "torch.cuda.synchronize()\n"
            "print('done', flush=True)\n")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, "-c", code], cwd=root, capture_output=True, text=True, timeout=300)
    assert "done" in r.stdout and r.returncode == 0, (r.returncode, r.stderr[-500:])


def test_views_outlive_close_and_keep_their_memory():
    """round-1 advisor finding: BatchEnv.close() destroyed the OkEnv under live DLPack views.  Now a tensor the caller
    still holds keeps the device memory alive: the env is destroyed when its last export is released."""
    import gc

    from openkitchen_b200 import dlpack

    env = ok.BatchEnv(["Monza"], 64, rays=15)
    env.step_random(5)
    torch.cuda.synchronize()
    keep = env.obs            # the caller's own reference
    want = keep.clone()
    inner = env.env
    env.close()
    assert env.obs is None and inner.h is not None, "the env must survive while a view is alive"
    assert torch.equal(keep, want)               # still valid memory
    other = ok.BatchEnv(["Spa"], 4096, rays=32)  # allocations in between must not land on it
    other.step_random(2)
    torch.cuda.synchronize()
    assert torch.equal(keep, want)
    other.close()
    del keep, want
    gc.collect()
    torch.cuda.synchronize()
    assert dlpack.live_exports(inner) == 0 and inner.h is None, "released with its last view"


def test_population_counters_are_the_buffer_sums():
    """ok_population_counters (per-step device counters) against the flag buffers themselves, with NVTX ranges on
    (OK_NVTX is read once per process: set here for whatever library calls come first in this process)."""
    import os

    os.environ.setdefault("OK_NVTX", "1")
    n = 5000
    env = ok.Env(device=0, auto_reset=0)
    for nm in ("Monza", "Spa"):
        env.add_named_track(nm)
    env.alloc_agents(n, ok.ray_fan(7), (np.arange(n) % 2).astype(np.int32))
    seen = []
    for k in range(6):
        env.launch_steps_random(8 * k, 8, 7)
        c = env.population_counters()
        seen.append(c["crashed"])
        crashed, timed_out, done = env.read("crashed"), env.read("timed_out"), env.read("done")
        assert c == {"agents": n, "alive": int((crashed == 0).sum()), "crashed": int((crashed != 0).sum()),
                     "timed_out": int((timed_out != 0).sum()), "done": int((done != 0).sum())}, (k, c)
        assert c["alive"] + c["crashed"] == n
    assert seen == sorted(seen) and seen[-1] > 0 and seen[0] < n, seen  # nobody is reset here: crashes only accumulate
