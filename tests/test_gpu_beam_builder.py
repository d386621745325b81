"""GPU: the device beam-table builder (openkitchen_b200/csrc/ok_beam_gpu.cu) produces tables with the same property the
host builder's tables have (tests/test_beam_cpu.py), and the kernels give bit-identical results with either."""
import numpy as np
import pytest

import openkitchen_b200 as ok
from tests.test_beam_cpu import check_beam_lists
from tests.util import assert_same, make_pair, spread_points

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name,cell,bins", [("Monza", 4.0, 128), ("Spa", 8.0, 64), ("Zandvoort", 2.0, 256)])
def test_device_built_lists_contain_the_reference_winner(monkeypatch, name, cell, bins):
    monkeypatch.setenv("OK_BEAM_BUILDER", "gpu")
    monkeypatch.setenv("OK_BEAM_CACHE", "0")
    check_beam_lists(ok.Env(device=0, beam_cell=cell, beam_bins=bins), name)


@pytest.mark.parametrize("builder", ["gpu", "cpu"])
def test_rollout_is_bit_exact_with_either_builder(monkeypatch, builder):
    monkeypatch.setenv("OK_BEAM_BUILDER", builder)
    monkeypatch.setenv("OK_BEAM_CACHE", "0")
    env, ora, tid = make_pair(["Austin", "Montreal", "Sepang"], 3 * 64, 32, raycast_mode=ok.RAYCAST_BEAM, beam_cell=3.0,
                              beam_bins=256, reward_mode=ok.REWARD_CMAES_PROGRESS, auto_reset=1)
    pts = spread_points(ora, tid)
    env.reset(None, pts)
    ora.reset(None, pts)
    for step in range(60):
        env.launch_steps_random(step, 1)
        ora.fill_random_actions(step)
        ora.step()
    assert_same(env, ora, ctx=f"{builder}-built tables")
