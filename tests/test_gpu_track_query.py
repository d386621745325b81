"""GPU: ok_track_query -- RaceTrack::findNearestTrackIndexBruteForce, getDistanceToLaneCenter and
getNearestDistanceToTrackBoundary (RaceTrack.cpp:16-72) on the device, bit-exact against the oracle and, when the
prebuilt checker is on the box, against the reference's own RaceTrack objects (oracle/_ref/libokref.so)."""
import numpy as np
import pytest

import openkitchen_b200 as ok
from oracle.api import Oracle, have_ref

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("kind", ["port", "reference"])
def test_track_queries_match_the_reference(kind):
    if kind == "reference" and not have_ref():
        pytest.skip("oracle/_ref/libokref.so not built")
    names = ["Monza", "Spa", "Oschersleben", "Austin"]
    env = ok.Env(device=0)
    ora = Oracle(kind)
    for nm in names:
        env.add_named_track(nm)
        ora.add_track(ok.track_columns(nm))
    rng = np.random.default_rng(7)
    n = 600
    tid = rng.integers(0, len(names), n).astype(np.int32)
    # points on the lane, far outside the window, and exactly on centre-line / boundary points (distance 0, ties)
    x = rng.uniform(0.0, 1600.0, n).astype(np.float32)
    y = rng.uniform(0.0, 1400.0, n).astype(np.float32)
    for k in range(0, n, 3):
        t = int(tid[k])
        cx, cy = env.track_array(t, "x"), env.track_array(t, "y")
        i = int(rng.integers(0, len(cx)))
        x[k], y[k] = cx[i] + np.float32(rng.normal(0, 6)), cy[i] + np.float32(rng.normal(0, 6))
    for k in range(1, 60, 3):
        t = int(tid[k])
        li = env.track_array(t, "li")
        i = len(li) - 1 if k < 10 else int(rng.integers(0, len(li)))  # the last point starts the closure segment
        x[k], y[k] = li[i]
    x[5], y[5] = np.float32(1e9), np.float32(-3e8)
    idx, lane, bound = env.track_query(x, y, tid)
    for k in range(n):
        t = int(tid[k])
        assert idx[k] == ora.nearest_index(t, float(x[k]), float(y[k])), k
        assert np.float32(lane[k]).tobytes() == np.float32(ora.dist_lane_center(t, float(x[k]), float(y[k]))).tobytes(), k
        assert np.float32(bound[k]).tobytes() == np.float32(ora.dist_boundary(t, float(x[k]), float(y[k]))).tobytes(), k
    assert (bound[1:60:3] == 0.0).all()


def test_track_query_on_agent_poses_equals_reward_mode_7():
    """the lane-centre reward the step kernel accumulates (OK_REWARD_LANE_CENTER) is 1 - the queried ratio"""
    env = ok.Env(device=0, reward_mode=ok.REWARD_LANE_CENTER)
    env.add_named_track("Sakhir")
    env.alloc_agents(64, ok.ray_fan(15))
    env.launch_steps_random(0, 20)
    x, y = env.read("pos_x"), env.read("pos_y")
    idx, lane, _ = env.track_query(x, y)
    alive = env.read("crashed") == 0
    assert alive.any()
    assert np.array_equal(idx[alive], env.read("nearest_idx")[alive])
    assert np.array_equal((np.float32(1.0) - lane)[alive].view(np.uint32), env.read("reward")[alive].view(np.uint32))
