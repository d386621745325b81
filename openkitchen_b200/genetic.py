"""The EvolutionaryRacer learner on the device.

* ``GeneticPopulation.act(env)``  GeneticAgent::updateAction + Network::infer (GeneticAgent.hpp:37-50,
  Network.hpp:119-155) for every agent in one kernel (``ok_genetic_policy``: one warp per agent, weights streamed once).
* ``GeneticPopulation.mate(scores)``  chooseAndMateAgents + mate2AgentsSelective (Mating.hpp:52-166): keep the top 5,
  slot 0 = clone of the best, slot 1 = self-mutation of the best, the rest = roulette-chosen distinct parents mated
  coefficient by coefficient (10 % random mutation in [-1,1], else 75 % the fitter parent's coefficient).  The reference
  draws from unseeded std::random_device generators; here a torch generator (same distributions, reproducible).
With torch.distributed the five parents are found on the gathered scores and their weights are broadcast from their owner.
"""
from __future__ import annotations

import torch
import torch.distributed as dist

from . import dist as okd

K_PARENTS = 5            # Mating.hpp:118
P_MUTATE = 0.1           # mate2AgentsSelective kMutationProb, Mating.hpp:54
P_DOMINANT = 0.75        # kDominantAgentThresh, Mating.hpp:55
HIDDEN, OUTPUTS = 30, 6  # Network.hpp:94-95


class GeneticPopulation:
    def __init__(self, n_agents: int, rays: int, device="cuda", generator: torch.Generator | None = None, hidden: int = HIDDEN):
        self.n, self.inputs, self.hidden = n_agents, rays + 2, hidden
        self.device = torch.device(device)
        self.gen = generator
        # Eigen ::Random(): uniform in [-1, 1] (Network.hpp:99-100)
        self.w1 = torch.rand(n_agents, self.inputs, hidden, device=self.device, generator=generator) * 2 - 1
        self.w2 = torch.rand(n_agents, hidden, OUTPUTS, device=self.device, generator=generator) * 2 - 1

    def act(self, env):
        """writes env.act_throttle / env.act_steer from env.speed / env.rot / env.obs"""
        env.env.genetic_policy(self.w1.data_ptr(), self.w2.data_ptr(), self.hidden, env._stream())

    def act_torch(self, env):
        """the same policy with library ops (bmm) -- the cross-check for the kernel"""
        rot = torch.remainder(env.rot, 360.0)
        x = torch.cat([(env.speed / 100.0)[:, None], (rot / 360.0)[:, None], env.obs], dim=1)
        h = torch.relu(torch.bmm(x[:, None, :], self.w1))
        z = torch.bmm(h, self.w2)[:, 0, :]
        on = torch.sigmoid(z) > 0.5
        thr = 0.3 * on[:, 0] - 0.3 * on[:, 1]
        steer = 1.0 * on[:, 2] + 4.0 * on[:, 3] - 1.0 * on[:, 4] - 4.0 * on[:, 5]
        return thr.float(), steer.float(), z

    # ---- mating --------------------------------------------------------------------------------
    def _mate_selective(self, dom_w, sub_w, draws=None):
        """mate2AgentsSelective (Mating.hpp:52-104) for a batch of offspring: dom_w / sub_w f32[count, ...] are the
        superior / inferior parent's weights.  Exactly the reference's two uniform draws per coefficient: r1 < 0.1 mutates
        to (r2 - 0.5) * 2, otherwise r2 < 0.75 takes the superior parent's coefficient.  draws: f32[count, ..., 2]
        (injected, for the parity test) or None (torch's generator)."""
        if draws is None:
            draws = torch.rand(*dom_w.shape, 2, device=self.device, generator=self.gen)
        r1, r2 = draws[..., 0], draws[..., 1]
        return torch.where(r1 < P_MUTATE, (r2 - 0.5) * 2.0, torch.where(r2 < P_DOMINANT, dom_w, sub_w))

    def _roulette(self, top_val, m):
        """(first, second != first) parents with probability proportional to the score (std::discrete_distribution with
        rejection of first == second, Mating.hpp:138-149 -- the same joint distribution, drawn without the loop)"""
        probs = top_val.clamp_min(0) + 1e-12
        first = torch.multinomial(probs.expand(max(m, 1), -1), 1, generator=self.gen)[:, 0]
        p2 = probs.expand(max(m, 1), -1).clone()
        p2.scatter_(1, first[:, None], 0.0)
        second = torch.multinomial(p2, 1, generator=self.gen)[:, 0]
        return first, second

    def mate_from_draws(self, pw1, pw2, top_val, first, second, draws_w1=None, draws_w2=None, with_elite=True):
        """chooseAndMateAgents (Mating.hpp:113-166) as a pure function of its random draws.  pw1 / pw2: the five parents'
        weights in score order, top_val their scores; first / second i64[m]: the roulette's parent indices for the m
        mated offspring; draws_w1 f32[k, inputs, hidden, 2], draws_w2 f32[k, hidden, 6, 2] with k = m (+ 1 leading row for
        the best agent's self-mutation when with_elite) or None.  Returns the new colony's (w1, w2): with_elite puts the
        best agent's clone in slot 0 and its self-mutation in slot 1 (Mating.hpp:127-131)."""
        m = int(first.numel())
        # "superior agent" = agent_1 only when its score is strictly greater (Mating.hpp:58-61): a tie favours agent_2
        first_wins = top_val[first] > top_val[second]
        dom, sub = torch.where(first_wins, first, second), torch.where(first_wins, second, first)
        off = 1 if with_elite else 0
        d1 = None if draws_w1 is None else draws_w1[off:off + m]
        d2 = None if draws_w2 is None else draws_w2[off:off + m]
        new1 = self._mate_selective(pw1[dom], pw1[sub], d1)
        new2 = self._mate_selective(pw2[dom], pw2[sub], d2)
        if not with_elite:
            return new1, new2
        s1 = self._mate_selective(pw1[:1], pw1[:1], None if draws_w1 is None else draws_w1[:1])
        s2 = self._mate_selective(pw2[:1], pw2[:1], None if draws_w2 is None else draws_w2[:1])
        return torch.cat([pw1[:1], s1, new1]), torch.cat([pw2[:1], s2, new2])

    def mate(self, scores: torch.Tensor):
        """scores f32[n_local] (higher is better, e.g. the nearest track index: MiscUtils.hpp:64-71).  Every rank must own
        a generator seeded differently (or the ranks breed identical offspring): pass e.g. seed * world + rank."""
        world = dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1
        rank = dist.get_rank() if world > 1 else 0
        n_total = self.n * world if world > 1 else self.n
        top_val, top_idx = okd.top_k(scores.to(self.device, torch.float32), K_PARENTS, n_total if world > 1 else None)
        if world > 1:
            lo, _ = okd.shard_bounds(n_total, rank, world)
            pw1 = torch.zeros(K_PARENTS, self.inputs, self.hidden, device=self.device)
            pw2 = torch.zeros(K_PARENTS, self.hidden, OUTPUTS, device=self.device)
            for k, gi in enumerate(top_idx.tolist()):
                if lo <= gi < lo + self.n:
                    pw1[k], pw2[k] = self.w1[gi - lo], self.w2[gi - lo]
            dist.all_reduce(pw1)
            dist.all_reduce(pw2)
        else:
            pw1, pw2 = self.w1[top_idx].clone(), self.w2[top_idx].clone()
        elite = rank == 0  # the global colony's slots 0 and 1 live on rank 0
        m = self.n - 2 if elite else self.n
        first, second = self._roulette(top_val, m)
        self.w1, self.w2 = self.mate_from_draws(pw1, pw2, top_val, first[:m], second[:m], with_elite=elite)
        return top_val, top_idx
