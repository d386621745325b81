#!/usr/bin/env python
"""CMA-ES racer on the batched environment -- the loop of CovarianceMatrixAdaptationEvolution/main_torch.cpp
(one candidate = one agent, throttle 100, steering = 5 * controller(lidar), fitness = sum |delta nearest index|,
zero if the agent timed out), with the whole population stepping in ONE kernel per tick and the solver update batched.

    python examples/cmaes_racer.py --population 4096 --generations 20
    torchrun --nproc-per-node 8 examples/cmaes_racer.py --population 8388608      # BASELINE config 4: 1M agents / GPU

Per generation the only cross-GPU traffic is the fitness all-gather and the all-reduce of the solver's partial sums."""
import argparse
import os
import sys
import time

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import openkitchen_b200 as ok  # noqa: E402
from openkitchen_b200.cmaes import CmaEs, PopulationController  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--population", type=int, default=4096)
    ap.add_argument("--generations", type=int, default=10)
    ap.add_argument("--ticks", type=int, default=600, help="episode cap (the reference runs until every agent crashed)")
    ap.add_argument("--rays", type=int, default=32)
    ap.add_argument("--track", default="Silverstone")
    args = ap.parse_args()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    rank = dist.get_rank() if world > 1 else 0
    ctrl = PopulationController(args.rays)
    # seed: the solver folds the rank in, so every rank draws DIFFERENT candidates for its slice of the population
    solver = CmaEs(ctrl.num_params, args.population, device=f"cuda:{local}", seed=1)
    n_local = solver.hi - solver.lo
    env = ok.BatchEnv([args.track], n_local, rays=args.rays, device=local, movement_mode=ok.MOVE_VELOCITY,
                      reward_mode=ok.REWARD_CMAES_PROGRESS)
    thr = torch.full((n_local,), 100.0, device="cuda")
    for g in range(args.generations):
        t0 = time.perf_counter()
        x = solver.sample()                       # f32[n_local, num_params]
        env.reset()                               # RaceTrack::kStartingIdx, as main_torch.cpp:119-123
        env.step(torch.zeros_like(thr), torch.zeros_like(thr))  # initial observation (zero action)
        ticks = 0
        for ticks in range(1, args.ticks + 1):
            ctrl.act(env, x)                      # one kernel: per-candidate MLP -> action buffers
            env.step()
            if ticks % 50 == 0 and bool(env.crashed.all()):
                break
        fitness, order = solver.tell(x, env.fitness)
        torch.cuda.synchronize()
        if rank == 0:
            dt = time.perf_counter() - t0
            print(f"generation {g}: best fitness {float(fitness[order[0]]):.0f}, mean {float(fitness.mean()):.1f}, "
                  f"sigma {solver.sigma:.3f}, {ticks} ticks, {args.population * ticks / dt:.3g} agent-steps/s", flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
