"""Rollout collection for the policy-gradient trainers on top of ``BatchEnv``.

The reference's PPO loop (RLRacers/PPO/ppo_sim.cpp:61-89, PPOAgent.hpp:68-102) builds one ``torch::tensor`` per agent
per tick, runs the actor on a batch of one, synchronises on ``.item<int>()`` for the sampled action and appends to
per-agent ``std::vector``s; the discounted returns are a host reverse scan (ExperienceBuffer.hpp:45-68).  Here the
whole rollout -- observation -> actor -> sample -> action table -> tick -> record, for every agent and ``steps`` ticks --
is captured once as ONE CUDA graph and replayed: no host work per tick, no launch latency between the ~10 small
kernels of a tick.  The returns are a batched reverse scan on the device.
"""
from __future__ import annotations

from typing import Callable

import torch

from .batch_env import BatchEnv


def discounted_returns(rewards: torch.Tensor, gamma: float = 0.99, done: torch.Tensor | None = None) -> torch.Tensor:
    """ExperienceBuffer::calculateDiscountedRewards (ExperienceBuffer.hpp:45-62) for every agent at once.

    rewards: f32[T, N] (one column per agent's ``saved_rewards``).  ret[t] = r[t] + gamma * ret[t + 1], scanned from the
    last tick backwards in binary32 with the reference's operation order.  ``done`` (u8/bool[T, N], optional) cuts the
    scan where an episode ended at tick t (the agent was reset afterwards), which the reference never needs because it
    rolls one episode per buffer."""
    if rewards.dim() != 2:
        raise ValueError("rewards must be [T, N]")
    out = torch.empty_like(rewards, dtype=torch.float32)
    acc = torch.zeros(rewards.shape[1], dtype=torch.float32, device=rewards.device)
    g = torch.tensor(gamma, dtype=torch.float32, device=rewards.device)
    for t in range(rewards.shape[0] - 1, -1, -1):
        if done is not None:
            acc = torch.where(done[t].bool(), torch.zeros_like(acc), acc)
        acc = rewards[t].to(torch.float32) + g * acc
        out[t] = acc
    return out


def normalize_returns(returns: torch.Tensor) -> torch.Tensor:
    """(x - mean) / (std + eps) per agent (ExperienceBuffer.hpp:64-67: unbiased std, eps = FLT_EPSILON)"""
    eps = torch.finfo(torch.float32).eps
    return (returns - returns.mean(dim=0, keepdim=True)) / (returns.std(dim=0, keepdim=True) + eps)


class GraphedRollout:
    """``steps`` ticks of every agent as one CUDA graph.

    policy: obs f32[N, R] -> action probabilities f32[N, A] (a ``torch.nn.Module`` or any CUDA-graph-safe callable;
    its parameters may be updated in place between replays).  action_table: f32[A, 2] = (throttle, steering) per
    discrete action (kActionMap of the RL racers).  sample=False takes the argmax instead of sampling.
    After ``run()``: ``obs`` f32[T, N, R] (the state each action was chosen from), ``actions`` i64[T, N], ``log_prob``
    f32[T, N], ``rewards`` f32[T, N], ``dones`` u8[T, N] (after the tick)."""

    def __init__(self, env: BatchEnv, policy: Callable[[torch.Tensor], torch.Tensor], action_table: torch.Tensor, steps: int,
                 sample: bool = True):
        self.env, self.policy, self.steps, self.sample = env, policy, int(steps), sample
        dev, n, r = env.device, env.n_agents, env.n_rays
        self.table = action_table.to(dev, torch.float32).contiguous()
        if self.table.dim() != 2 or self.table.shape[1] != 2:
            raise ValueError("action_table must be [A, 2] = (throttle, steering)")
        self.obs = torch.empty(self.steps, n, r, device=dev)
        self.actions = torch.empty(self.steps, n, dtype=torch.int64, device=dev)
        self.log_prob = torch.empty(self.steps, n, device=dev)
        self.rewards = torch.empty(self.steps, n, device=dev)
        self.dones = torch.empty(self.steps, n, dtype=torch.uint8, device=dev)
        self._thr = torch.empty(n, device=dev)
        self._steer = torch.empty(n, device=dev)
        self.graph: torch.cuda.CUDAGraph | None = None

    @torch.no_grad()
    def _tick(self, t: int):
        obs = self.env.obs
        self.obs[t].copy_(obs)
        probs = self.policy(obs)
        a = torch.multinomial(probs, 1)[:, 0] if self.sample else probs.argmax(dim=-1)
        self.actions[t].copy_(a)
        self.log_prob[t].copy_(torch.log(probs.gather(1, a[:, None])[:, 0]))
        torch.index_select(self.table[:, 0], 0, a, out=self._thr)
        torch.index_select(self.table[:, 1], 0, a, out=self._steer)
        self.env.step(self._thr, self._steer)
        self.rewards[t].copy_(self.env.reward)
        self.dones[t].copy_(self.env.done)

    def run_eager(self):
        """the same rollout launched tick by tick (what the graph replaces; also the capture warm-up)"""
        for t in range(self.steps):
            self._tick(t)
        return self

    def capture(self):
        dev = self.env.device
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):
            self._tick(0)  # warm-up outside the capture: lazy initialisation of the policy / cuBLAS workspaces
            torch.cuda.synchronize(dev)
            self.graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(self.graph, stream=side):
                self.run_eager()
        torch.cuda.current_stream(dev).wait_stream(side)
        torch.cuda.synchronize(dev)
        return self

    def run(self):
        """one rollout of ``steps`` ticks from the env's current state (captures the graph on first use)"""
        if self.graph is None:
            self.capture()
        self.graph.replay()
        return self
