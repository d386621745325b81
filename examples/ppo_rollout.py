#!/usr/bin/env python
"""PPO-style rollout collection on the batched environment -- the loop of RLRacers/PPO/ppo_sim.cpp:61-89 with the
reference's Actor (Actor.hpp:9-26: Linear(5,128) -> relu -> Linear(128,3) -> softmax) evaluated, sampled and applied by
ONE kernel per tick (ok_ppo_actor), all ticks of a rollout replayed as one CUDA graph, returns by the discounted-return
kernel.  The policy update (PPOAgent.hpp:104-154) stays in torch on the recorded tensors.

    python examples/ppo_rollout.py --envs 4096 --steps 256 --iterations 20
"""
import argparse
import os
import sys
import time

import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import openkitchen_b200 as ok  # noqa: E402
from openkitchen_b200.rollout import FusedActorRollout, discounted_returns_fused, normalize_returns  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--envs", type=int, default=4096)
    ap.add_argument("--steps", type=int, default=256)
    ap.add_argument("--iterations", type=int, default=10)
    ap.add_argument("--track", default="Monza")
    args = ap.parse_args()
    fan = [-70.0, -30.0, 0.0, 30.0, 70.0]                                   # PPOAgent.hpp:57-61
    table = torch.tensor([[60.0, 0.0], [30.0, 5.0], [30.0, -5.0]])          # PPOAgent::kActionMap
    env = ok.BatchEnv([args.track], args.envs, rays=fan, reward_mode=ok.REWARD_CONSTANT, auto_reset=1)
    l1, l2 = torch.nn.Linear(5, 128).cuda(), torch.nn.Linear(128, 3).cuda()  # Actor::l1 / l2
    critic = torch.nn.Sequential(torch.nn.Linear(5, 128), torch.nn.ReLU(), torch.nn.Linear(128, 1)).cuda()
    opt = torch.optim.Adam(list(l1.parameters()) + list(l2.parameters()) + list(critic.parameters()), lr=3e-4)
    ro = FusedActorRollout(env, l1, l2, table, steps=args.steps)
    env.reset_random()
    env.cast_rays()
    for it in range(args.iterations):
        t0 = time.perf_counter()
        ro._tick0 = it * args.steps  # a fresh slice of the Philox stream per rollout
        ro.run_eager() if it == 0 else ro.run()
        ret = normalize_returns(discounted_returns_fused(env, ro.rewards, 0.99, ro.dones))
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        # one clipped-surrogate epoch over the whole buffer (PPOAgent.hpp:122-146)
        obs, act, old = ro.obs.reshape(-1, 5), ro.actions.reshape(-1).long(), ro.log_prob.reshape(-1)
        probs = torch.softmax(l2(torch.relu(l1(obs))), dim=1).clamp(1e-8, 1 - 1e-8)
        new = torch.log(probs.gather(1, act[:, None])[:, 0])
        value = critic(obs)[:, 0]
        adv = ret.reshape(-1) - value.detach()
        ratio = torch.exp(new - old)
        loss = -torch.min(ratio * adv, ratio.clamp(0.8, 1.2) * adv).mean() + torch.nn.functional.mse_loss(value, ret.reshape(-1))
        opt.zero_grad()
        loss.backward()
        opt.step()
        print(f"iteration {it}: {args.envs * args.steps / dt:.3g} env-steps/s collected, crash rate {float(ro.dones.float().mean()):.4f}, "
              f"loss {float(loss):.4f}", flush=True)


if __name__ == "__main__":
    main()
