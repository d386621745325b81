#!/usr/bin/env python
"""Measure the other BASELINE.json configs on one B200 (bench.py measures configs[2], the headline shape).

  C1  single agent on Monza, 5-ray Q-learning fan, VELOCITY mode, random actions
  C2  4,096-agent EvolutionaryRacer population on Silverstone, 32 rays, ACCELERATION mode, per-agent
      shallow MLP (R+2 -> 30 relu -> 6 sigmoid, GeneticAgent.hpp:37-50 action decoding) evaluated with torch.bmm
      on the zero-copy DLPack views
  C4  PPO rollout collection: 4,096 envs x 256 steps = 1,048,576 env-steps, shared torch actor 5->128->3
      (PPO/Actor.hpp:9-26) sampling actions on the device
  C5s one GPU's shard of the 8M-agent CMA-ES population: 1,048,576 agents, 32 rays, random actions
Each line: agent-steps/s, ray-casts/s, ms per tick (CUDA events, 20 warm-up ticks).
"""
import json
import sys
import time

import numpy as np
import torch

sys.path.insert(0, ".")
import openkitchen_b200 as ok  # noqa: E402


def timed(fn, ticks, warmup=20):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(ticks):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / ticks


def report(name, n, rays, ms, extra=None):
    line = {"config": name, "agents": n, "rays": rays, "ms_per_tick": ms, "agent_steps_per_sec": n / (ms * 1e-3),
            "ray_casts_per_sec": n * rays / (ms * 1e-3)}
    line.update(extra or {})
    print(json.dumps(line), flush=True)


def c1():
    env = ok.BatchEnv(["Monza"], 1, rays=[-70, -30, 0, 30, 70], reward_mode=ok.REWARD_Q_PROGRESS, auto_reset=1)
    ms = timed(lambda: env.step_random(1), 2000, 100)
    report("C1 single agent, Monza, 5 rays", 1, 5, ms, {"note": "latency of one tick: launch + one CTA"})
    g = torch.cuda.CUDAGraph()
    env.step_random(1)
    torch.cuda.synchronize()
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        env.step_random(1)
        torch.cuda.synchronize()
        with torch.cuda.graph(g, stream=s):
            env.env.launch_steps_random(1000, 100, 0x0C17C4E2, torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    ms = timed(g.replay, 20, 3) / 100
    report("C1 single agent, Monza, 5 rays (100 ticks per CUDA graph)", 1, 5, ms)


def c2():
    n, rays = 4096, 32
    env = ok.BatchEnv(["Silverstone"], n, rays=rays, movement_mode=ok.MOVE_ACCELERATION, reward_mode=ok.REWARD_TRACK_INDEX)
    g = torch.Generator(device="cuda").manual_seed(0)
    w1 = torch.rand(n, rays + 2, 30, device="cuda", generator=g) * 2 - 1  # Eigen ::Random(), Network.hpp:99-100
    w2 = torch.rand(n, 30, 6, device="cuda", generator=g) * 2 - 1
    env.reset_random(generator=g)
    env.cast_rays()

    def tick():
        # Network::infer, Network.hpp:119-155: inputs speed/100, normalizeAngleDeg(rot)/360, hits/200
        x = torch.cat([(env.speed / 100.0)[:, None], (torch.remainder(env.rot, 360.0) / 360.0)[:, None], env.obs], dim=1)
        h = torch.relu(torch.bmm(x[:, None, :], w1))
        z = torch.bmm(h, w2)[:, 0, :]
        on = z > 0  # sigmoid(z) > 0.5
        thr = 0.3 * on[:, 0] - 0.3 * on[:, 1]
        steer = 1.0 * on[:, 2] + 4.0 * on[:, 3] - 1.0 * on[:, 4] - 4.0 * on[:, 5]
        env.step(thr.float(), steer.float())

    ms = timed(tick, 300)
    ms_env = timed(lambda: env.step(), 300)
    report("C2 4096-agent genetic population, Silverstone, 32 rays, per-agent MLP (torch.bmm) + tick", n, rays, ms,
           {"ms_env_only": ms_env})
    from openkitchen_b200.genetic import GeneticPopulation

    pop = GeneticPopulation(n, rays, generator=g)

    def tick2():
        pop.act(env)   # ok_genetic_policy: one warp per agent, weights streamed once
        env.step()

    ms2 = timed(tick2, 300)
    ms_pol = timed(lambda: pop.act(env), 300)
    report("C2 4096-agent genetic population, Silverstone, 32 rays, per-agent MLP (ok_genetic_policy kernel) + tick", n, rays, ms2,
           {"ms_policy_only": ms_pol, "policy_weight_bytes": n * (rays + 2 + 6) * 30 * 4})
    # the policy kernel at a size where HBM, not launch latency, is the bound
    nb = 262144
    big = ok.BatchEnv(["Silverstone"], nb, rays=rays, movement_mode=ok.MOVE_ACCELERATION)
    pb = GeneticPopulation(nb, rays, generator=g)
    big.cast_rays()
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device="cuda")
    evs = []
    for _ in range(20):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); pb.act(big); e1.record()
        evs.append((e0, e1))
    torch.cuda.synchronize()
    msb = float(np.median([a.elapsed_time(b) for a, b in evs[3:]]))
    bytes_ = nb * ((rays + 2 + 6) * 30 * 4 + rays * 4 + 8 + 8)
    report("policy kernel alone, 262,144 agents (L2 flushed)", nb, rays, msb,
           {"algorithmic_bytes": bytes_, "GBps": bytes_ / (msb * 1e-3) / 1e9, "frac_of_measured_hbm_6553": bytes_ / (msb * 1e-3) / 1e9 / 6553.3})


def c4():
    n, rays, steps = 4096, 5, 256
    env = ok.BatchEnv(["Monza"], n, rays=[-70, -30, 0, 30, 70], reward_mode=ok.REWARD_CONSTANT)
    actor = torch.nn.Sequential(torch.nn.Linear(5, 128), torch.nn.ReLU(), torch.nn.Linear(128, 3), torch.nn.Softmax(-1)).cuda()
    amap_thr = torch.tensor([60.0, 30.0, 30.0], device="cuda")  # kActionMap-style (throttle, steer) table
    amap_st = torch.tensor([0.0, 4.0, -4.0], device="cuda")
    obs_buf = torch.empty(steps, n, 5, device="cuda")
    act_buf = torch.empty(steps, n, dtype=torch.int64, device="cuda")
    rew_buf = torch.empty(steps, n, device="cuda")
    done_buf = torch.empty(steps, n, dtype=torch.uint8, device="cuda")

    @torch.no_grad()
    def rollout():
        env.reset_random()
        env.cast_rays()
        for t in range(steps):
            p = actor(env.obs).clamp(1e-6, 1 - 1e-6)
            a = torch.multinomial(p, 1)[:, 0]
            obs_buf[t].copy_(env.obs)
            act_buf[t] = a
            env.step(amap_thr[a], amap_st[a])
            rew_buf[t].copy_(env.reward)
            done_buf[t].copy_(env.done)
            env.reset_random(env.done.bool())

    rollout()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    rollout()
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    report("C4 PPO rollout 4096 envs x 256 steps, torch actor 5-128-3 on DLPack views", n, rays, 1e3 * dt / steps,
           {"env_steps": n * steps, "env_steps_per_sec": n * steps / dt, "wall_s": dt})

    # the same collection as ONE CUDA graph (openkitchen_b200.rollout.GraphedRollout): in-kernel auto-reset instead of
    # the host-synchronising reset of done agents, no host work between the ticks
    from openkitchen_b200.rollout import GraphedRollout, discounted_returns

    env2 = ok.BatchEnv(["Monza"], n, rays=[-70, -30, 0, 30, 70], reward_mode=ok.REWARD_CONSTANT, auto_reset=1)
    table = torch.stack([amap_thr, amap_st], dim=1)
    clamped = torch.nn.Sequential(actor, torch.nn.Hardtanh(1e-6, 1 - 1e-6))
    ro = GraphedRollout(env2, clamped, table, steps, sample=True)
    env2.reset_random()
    env2.cast_rays()
    ro.capture()
    ro.run()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    reps = 5
    for _ in range(reps):
        ro.run()
        ret = discounted_returns(ro.rewards, 0.99, ro.dones)
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / reps
    report("C4 PPO rollout 4096 envs x 256 steps as one CUDA graph (GraphedRollout) + device return scan", n, rays,
           1e3 * dt / steps, {"env_steps": n * steps, "env_steps_per_sec": n * steps / dt, "wall_s": dt,
                              "crashed_fraction": float(ro.dones.float().mean())})


    # the policy step as phase 0 of the step kernel's tiles (ok_ppo_actor_step): ONE launch per tick, all ticks in one graph,
    # returns by the discounted-return kernel (openkitchen_b200.rollout.FusedActorRollout)
    from openkitchen_b200.rollout import FusedActorRollout, discounted_returns_fused

    env3 = ok.BatchEnv(["Monza"], n, rays=[-70, -30, 0, 30, 70], reward_mode=ok.REWARD_CONSTANT, auto_reset=1)
    fr = FusedActorRollout(env3, actor[0], actor[2], table, steps, sample=True)
    env3.reset_random()
    env3.cast_rays()
    fr.capture()
    fr.run()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record()
    for i in range(reps):
        fr.run()
        ret = discounted_returns_fused(env3, fr.rewards, 0.99, fr.dones)
    e1.record()
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / reps
    report("C4 PPO rollout 4096 envs x 256 steps, policy step fused into the step kernel: one launch per tick, one CUDA graph (FusedActorRollout) + return kernel",
           n, rays, 1e3 * dt / steps, {"env_steps": n * steps, "env_steps_per_sec": n * steps / dt, "wall_s": dt,
                                       "device_ms_per_tick": e0.elapsed_time(e1) / reps / steps,
                                       "crashed_fraction": float(fr.dones.float().mean())})


def c5s():
    n, rays = 1_048_576, 32
    env = ok.BatchEnv(ok.track_names(), n, rays=rays, reward_mode=ok.REWARD_CMAES_PROGRESS, auto_reset=1)
    ids = np.arange(n, dtype=np.uint64)
    pts_per = np.asarray(env.points_per_track, dtype=np.uint64)[env.track_id.numpy()]
    env.reset(None, torch.as_tensor(((ids * np.uint64(2654435761)) % np.uint64(2**32) % pts_per).astype(np.int32)))
    ms = timed(lambda: env.step_random(1), 50, 10)
    report("C5 shard: 1,048,576 agents (1/8 of the 8M CMA-ES population), 23 tracks, 32 rays", n, rays, ms)


if __name__ == "__main__":
    which = sys.argv[1:] or ["c1", "c2", "c4", "c5s"]
    for w in which:
        {"c1": c1, "c2": c2, "c4": c4, "c5s": c5s}[w]()
