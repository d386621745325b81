"""GPU: the per-agent shallow-MLP policy kernel (ok_genetic_policy) against a float64 numpy evaluation of
Network::infer + GeneticAgent::updateAction, and the batched mating rules."""
import numpy as np
import pytest

import openkitchen_b200 as ok

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")


def _reference_actions(w1, w2, speed, rot, obs):
    """Network.hpp:119-155 + GeneticAgent.hpp:37-50, one agent at a time, float64 accumulation; returns the actions
    and the pre-activations z so that decisions with |z| tiny (rounding-order dependent) can be excused"""
    n = len(speed)
    thr, st, zs = np.zeros(n, np.float32), np.zeros(n, np.float32), np.zeros((n, 6))
    for a in range(n):
        r = float(rot[a])
        while r < 360.0:
            r += 360.0
        while r >= 360.0:
            r -= 360.0
        x = np.concatenate([[speed[a] / 100.0, r / 360.0], obs[a]]).astype(np.float64)
        h = np.maximum(x @ w1[a].astype(np.float64), 0.0)
        z = h @ w2[a].astype(np.float64)
        on = z > 0
        thr[a] = 0.3 * on[0] - 0.3 * on[1]
        st[a] = 1.0 * on[2] + 4.0 * on[3] - 1.0 * on[4] - 4.0 * on[5]
        zs[a] = z
    return thr, st, zs


@pytest.mark.parametrize("rays", [15, 32, 40])
def test_policy_kernel_matches_reference_network(rays):
    from openkitchen_b200.genetic import GeneticPopulation

    n = 512
    env = ok.BatchEnv(["Silverstone"], n, rays=rays, movement_mode=ok.MOVE_ACCELERATION)
    g = torch.Generator(device="cuda").manual_seed(3)
    pop = GeneticPopulation(n, rays, generator=g)
    env.reset_random(generator=g)
    env.cast_rays()
    env.rot.add_(torch.randint(-3, 4, (n,), device="cuda", generator=g).float() * 360.0)  # exercise normalizeAngleDeg
    for _ in range(5):
        pop.act(env)
        torch.cuda.synchronize()
        thr, st, z = _reference_actions(pop.w1.cpu().numpy(), pop.w2.cpu().numpy(), env.speed.cpu().numpy(),
                                        env.rot.cpu().numpy(), env.obs.cpu().numpy())
        sure = (np.abs(z) > 1e-4).all(axis=1)  # decisions that do not hinge on summation order
        assert sure.mean() > 0.99
        assert np.array_equal(env.act_throttle.cpu().numpy()[sure], thr[sure])
        assert np.array_equal(env.act_steer.cpu().numpy()[sure], st[sure])
        t2, s2, _ = pop.act_torch(env)
        assert (t2.cpu().numpy()[sure] == thr[sure]).all() and (s2.cpu().numpy()[sure] == st[sure]).all()
        env.step()


def test_mating_rules():
    from openkitchen_b200.genetic import GeneticPopulation

    n, rays = 200, 15
    g = torch.Generator(device="cuda").manual_seed(0)
    pop = GeneticPopulation(n, rays, generator=g)
    scores = torch.arange(n, device="cuda", dtype=torch.float32)
    old1 = pop.w1.clone()
    top_val, top_idx = pop.mate(scores)
    assert top_idx.tolist() == [199, 198, 197, 196, 195]
    parents = old1[top_idx]
    assert torch.equal(pop.w1[0], parents[0])  # the clone of the best
    # every offspring coefficient is one of the five parents' coefficients or a fresh value in [-1, 1]
    from_parent = (pop.w1[:, None] == parents[None]).any(1)
    assert (pop.w1.abs() <= 1).all()
    frac_mut = 1.0 - from_parent[2:].float().mean().item()
    assert 0.07 < frac_mut < 0.13  # kMutationProb = 0.1
    self_mut = pop.w1[1]
    assert 0.8 < (self_mut == parents[0]).float().mean().item() < 0.97
    assert pop.w1.shape == old1.shape and pop.w2.shape[0] == n
