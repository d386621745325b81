"""CPU: the batched discounted-return scan against a per-agent restatement of ExperienceBuffer::calculateDiscountedRewards
(RLRacers/PPO/ExperienceBuffer.hpp:45-67)."""
import numpy as np
import torch

from openkitchen_b200.rollout import discounted_returns, normalize_returns


def _reference_returns(saved_rewards, gamma=np.float32(0.99)):
    """one agent's buffer, the reference's loop in binary32"""
    out = np.zeros(len(saved_rewards), dtype=np.float32)
    cum = np.float32(0.0)
    for i in range(len(saved_rewards) - 1, -1, -1):
        cum = np.float32(saved_rewards[i] + np.float32(gamma * cum))
        out[i] = cum
    return out


def test_discounted_returns_match_the_reference_loop_bit_for_bit():
    rng = np.random.default_rng(0)
    r = rng.normal(size=(200, 17)).astype(np.float32)
    r[:, 0] = 1.0  # ppo_sim.cpp:76: constant reward
    got = discounted_returns(torch.from_numpy(r)).numpy()
    for a in range(r.shape[1]):
        want = _reference_returns(r[:, a])
        assert np.array_equal(got[:, a].view(np.uint32), want.view(np.uint32)), f"agent {a}"


def test_done_cuts_the_scan_at_episode_ends():
    r = torch.ones(6, 2)
    done = torch.zeros(6, 2, dtype=torch.uint8)
    done[2, 0] = 1  # agent 0's episode ends with tick 2
    got = discounted_returns(r, gamma=0.5, done=done)
    assert torch.allclose(got[:, 0], torch.tensor([1.75, 1.5, 1.0, 1.75, 1.5, 1.0]))
    assert torch.allclose(got[:, 1], torch.tensor([1.96875, 1.9375, 1.875, 1.75, 1.5, 1.0]))


def test_normalize_returns_is_the_reference_formula():
    rng = np.random.default_rng(1)
    x = torch.from_numpy(rng.normal(size=(50, 3)).astype(np.float32))
    got = normalize_returns(x)
    for a in range(3):
        col = x[:, a]
        want = (col - col.mean()) / (col.std() + torch.finfo(torch.float32).eps)
        assert torch.allclose(got[:, a], want, rtol=1e-6, atol=1e-6)
