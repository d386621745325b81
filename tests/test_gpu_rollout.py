"""GPU: the CUDA-graph rollout (openkitchen_b200.rollout.GraphedRollout) records exactly what the same
loop launched tick by tick records."""
import numpy as np
import pytest
import torch

import openkitchen_b200 as ok
from openkitchen_b200.rollout import GraphedRollout, discounted_returns

pytestmark = pytest.mark.gpu


def _make(n):
    env = ok.BatchEnv(["Monza"], n, rays=[-70, -30, 0, 30, 70], reward_mode=ok.REWARD_CONSTANT, auto_reset=1)
    torch.manual_seed(0)
    actor = torch.nn.Sequential(torch.nn.Linear(5, 32), torch.nn.ReLU(), torch.nn.Linear(32, 3), torch.nn.Softmax(-1)).cuda()
    table = torch.tensor([[60.0, 0.0], [30.0, 4.0], [30.0, -4.0]], device="cuda")
    return env, actor, table


def test_graph_replay_equals_eager_rollout():
    n, steps = 512, 40
    env_a, actor, table = _make(n)
    env_b, _, _ = _make(n)
    ra = GraphedRollout(env_a, actor, table, steps, sample=False).capture()
    rb = GraphedRollout(env_b, actor, table, steps, sample=False)
    pts = torch.arange(n, device="cuda", dtype=torch.int32) * 7 % 1000
    for env in (env_a, env_b):
        env.reset(None, pts)
        env.cast_rays()
    for rep in range(2):  # the second rollout continues from where the first one stopped
        ra.run()
        rb.run_eager()
        torch.cuda.synchronize()
        for name in ("obs", "actions", "log_prob", "rewards", "dones"):
            assert torch.equal(getattr(ra, name), getattr(rb, name)), f"{name}, rollout {rep}"
        for name in ("pos_x", "pos_y", "rot", "crashed"):
            assert torch.equal(env_a[name], env_b[name]), f"{name}, rollout {rep}"
    assert ra.dones.sum() > 0, "nobody crashed: the rollout is too short to test resets"
    ret = discounted_returns(ra.rewards, 0.99, ra.dones)
    assert ret.shape == (steps, n) and torch.isfinite(ret).all() and (ret[-1] == 1.0).all()


def test_sampled_rollout_is_consistent():
    n, steps = 256, 16
    env, actor, table = _make(n)
    r = GraphedRollout(env, actor, table, steps, sample=True)
    env.reset(None, torch.full((n,), 3, device="cuda", dtype=torch.int32))
    env.cast_rays()
    r.run()
    torch.cuda.synchronize()
    assert ((r.actions >= 0) & (r.actions < 3)).all()
    p = actor(r.obs.reshape(-1, 5)).reshape(steps, n, 3)
    assert torch.allclose(r.log_prob, torch.log(p.gather(2, r.actions[..., None])[..., 0]), rtol=1e-5, atol=1e-6)
    assert len(torch.unique(r.actions)) == 3
