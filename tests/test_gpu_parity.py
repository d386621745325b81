"""GPU parity proper: the sm_100a path, called through the C ABI, against the CPU oracle on the
same seeded inputs.  Bar: BIT-EXACT on every buffer (flags, indices AND the binary32 poses /
distances -- stricter than the 1e-4 relative the north star asks for floats)."""
import numpy as np
import pytest

import openkitchen_b200 as ok
from oracle.api import have_ref
from tests.util import ALL_BUFS, assert_same, make_pair, spread_points

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("mode", [ok.MOVE_VELOCITY, ok.MOVE_ACCELERATION])
@pytest.mark.parametrize("raycast,kind", [(ok.RAYCAST_BEAM, "port"), (ok.RAYCAST_GRID, "port"), (ok.RAYCAST_BRUTE, "port"),
                                          (ok.RAYCAST_BEAM, "reference")])
def test_rollout_matches_oracle(mode, raycast, kind):
    """kind = "reference": the checker is oracle/_ref/libokref.so, i.e. the reference's own Agent.cpp / RaceTrack.cpp
    objects (shipped prebuilt to the GPU box), not the C restatement."""
    if kind == "reference" and not have_ref():
        pytest.skip("oracle/_ref/libokref.so not built")
    names = ok.track_names()
    env, ora, tid = make_pair(names, 23 * 12, 32, kind=kind, movement_mode=mode, raycast_mode=raycast,
                              reward_mode=ok.REWARD_CMAES_PROGRESS, auto_reset=1)
    pts = spread_points(ora, tid)
    env.reset(None, pts)
    ora.reset(None, pts)
    assert_same(env, ora, ctx="after reset")
    for step in range(120):
        env.launch_steps_random(step, 1)
        ora.fill_random_actions(step)
        ora.step()
        if step % 20 == 0 or step == 119:
            assert_same(env, ora, ctx=f"step {step}")
    if mode == ok.MOVE_VELOCITY:
        assert (ora.buffer("reset_pt") != pts).sum() > 0, "rollout never crashed and auto-reset: test too weak"


@pytest.mark.parametrize("reward", [1, 2, 3, 4, 5, 6, 7])
def test_reward_modes(reward):
    env, ora, tid = make_pair(["Monza", "Spa"], 64, 15, reward_mode=reward, auto_reset=0)
    pts = spread_points(ora, tid)
    env.reset(None, pts)
    ora.reset(None, pts)
    for step in range(60):
        env.launch_steps_random(step, 1)
        ora.fill_random_actions(step)
        ora.step()
    assert_same(env, ora, ctx=f"reward mode {reward}")


@pytest.mark.parametrize("rays", [1, 5, 7, 15, 32, 128])
def test_ray_counts(rays):
    fan = np.array([-70, -30, 0, 30, 70], dtype=np.float32) if rays == 5 else rays
    env, ora, tid = make_pair(["Monza"], 37, fan, reward_mode=ok.REWARD_MIN_RAY, auto_reset=1)
    pts = spread_points(ora, tid)
    env.reset(None, pts)
    ora.reset(None, pts)
    for step in range(40):
        env.launch_steps_random(step, 1)
        ora.fill_random_actions(step)
        ora.step()
    assert_same(env, ora, ctx=f"{rays} rays")


def test_single_agent_monza_q_learning_setup():
    """BASELINE config 0: one agent on Monza, 5-ray Q-learning fan, VELOCITY mode, 1000 ticks."""
    fan = np.array([-70, -30, 0, 30, 70], dtype=np.float32)
    env, ora, _ = make_pair(["Monza"], 1, fan, reward_mode=ok.REWARD_Q_PROGRESS, auto_reset=1)
    for step in range(1000):
        env.launch_steps_random(step, 1)
        ora.fill_random_actions(step)
        ora.step()
    assert_same(env, ora, ctx="C1 1000 ticks")


def test_external_actions_and_host_step():
    env, ora, tid = make_pair(["Silverstone"], 50, 32, movement_mode=ok.MOVE_ACCELERATION, reward_mode=ok.REWARD_CONSTANT)
    rng = np.random.default_rng(0)
    obs = np.zeros((50, 32), dtype=np.float32)
    rew = np.zeros(50, dtype=np.float32)
    done = np.zeros(50, dtype=np.uint8)
    for step in range(30):
        thr = rng.choice(np.array([-0.3, 0.0, 0.3], dtype=np.float32), 50)
        steer = rng.choice(np.array([-4, -1, 0, 1, 4], dtype=np.float32), 50)
        env.step_host(thr, steer, obs, rew, done)
        ora.step(thr, steer)
        assert np.array_equal(obs.view(np.uint32), ora.buffer("obs").view(np.uint32))
        assert np.array_equal(rew, ora.buffer("reward")) and np.array_equal(done, ora.buffer("done"))
    assert_same(env, ora, ctx="host step")


def test_reset_variants_and_cast_rays():
    env, ora, tid = make_pair(["Austin", "Zandvoort"], 40, 15)
    rng = np.random.default_rng(1)
    idx = rng.permutation(40)[:25].astype(np.int64)
    pts = np.array([rng.integers(0, ora.track_points(int(tid[i]))) for i in idx], dtype=np.int32)
    alpha = (rng.integers(10, 91, size=25) / np.float32(100.0)).astype(np.float32)
    hoff = np.where(np.arange(25) % 2 == 0, -1.0, 1.0).astype(np.float32) * (45 + rng.integers(0, 46, size=25)).astype(np.float32)
    env.reset(idx, pts, alpha, hoff)
    ora.reset(idx, pts, alpha, hoff)
    assert_same(env, ora, ctx="lane/heading reset")
    env.cast_rays()
    ora.cast_rays()
    assert_same(env, ora, ctx="cast_rays")
    # last centre-line point: the lane lerp must read the closure segments
    last = np.array([ora.track_points(int(tid[i])) - 1 for i in idx], dtype=np.int32)
    env.reset(idx, last, alpha, None)
    ora.reset(idx, last, alpha, None)
    env.cast_rays()
    ora.cast_rays()
    assert_same(env, ora, ctx="reset at last point")


def test_standstill_timeout_and_stale_hits():
    """zero actions: nobody moves, everyone times out on tick 201 and keeps stale hits afterwards."""
    env, ora, tid = make_pair(["Monza"], 8, 15, reward_mode=ok.REWARD_CMAES_PROGRESS)
    z = np.zeros(8, dtype=np.float32)
    for step in range(205):
        env.step_host(z, z)
        ora.step(z, z)
        if step in (0, 199, 200, 201, 204):
            assert_same(env, ora, ctx=f"standstill step {step}")
    assert ora.buffer("timed_out").all()


def test_far_away_and_nonfinite_poses():
    """poses outside the grid, huge headings (large-argument sincos path) and NaN/inf positions"""
    env, ora, tid = make_pair(["Monza"], 12, 32)
    x = np.array([-500, 5000, 800, 800, 1e6, -1e6, 800, np.nan, np.inf, 44, 1544, 800], dtype=np.float32)
    y = np.array([700, 700, -300, 4000, 1e6, 3, 700, 700, 700, 44, 1355, np.nan], dtype=np.float32)
    rot = np.array([0, 180, 90, -90, 225, 0, 1e9, 10, 20, 45, 225, 0], dtype=np.float32)
    for e in (env,):
        e.write("pos_x", x), e.write("pos_y", y), e.write("rot", rot)
    ora.buffer("pos_x")[:] = x
    ora.buffer("pos_y")[:] = y
    ora.buffer("rot")[:] = rot
    env.cast_rays()
    ora.cast_rays()
    assert_same(env, ora, names=["hit_abs", "hit_rel", "hit_seg", "hit_t", "min_dist2", "crashed", "obs"], ctx="far poses")


def test_interleaved_tracks():
    """agents of different tracks interleaved (worst case for the tile table) still match"""
    env, ora, tid = make_pair(["Monza", "Spa", "IMS"], 45, 7, grouped=False, auto_reset=1, reward_mode=2)
    pts = spread_points(ora, tid)
    env.reset(None, pts)
    ora.reset(None, pts)
    for step in range(30):
        env.launch_steps_random(step, 1)
        ora.fill_random_actions(step)
        ora.step()
    assert_same(env, ora, ctx="interleaved")



@pytest.mark.parametrize("rays", [5, 32, 48])
def test_many_tracks_per_cta_beam(rays):
    """several tracks and small tiles per CTA (track re-staging between batches), ray counts that do and do not fill
    warps, teleported and non-finite poses in the middle of a rollout (the windowed nearest-index search must fall
    back to the full search for them)"""
    names = ok.track_names()[:5]
    env, ora, tid = make_pair(names, 5 * 97, rays, raycast_mode=ok.RAYCAST_BEAM,
                              reward_mode=ok.REWARD_CMAES_PROGRESS, auto_reset=1)
    pts = spread_points(ora, tid)
    env.reset(None, pts)
    ora.reset(None, pts)
    rng = np.random.default_rng(3)
    for step in range(60):
        if step == 20:
            x, y = ora.buffer("pos_x").copy(), ora.buffer("pos_y").copy()
            x[::7] += rng.uniform(-300, 300, size=x[::7].shape).astype(np.float32)
            y[::7] += rng.uniform(-300, 300, size=y[::7].shape).astype(np.float32)
            x[1], y[1] = np.float32(np.nan), np.float32(5.0)
            x[2], y[2] = np.float32(1e30), np.float32(-1e30)
            env.write("pos_x", x), env.write("pos_y", y)
            ora.buffer("pos_x")[:] = x
            ora.buffer("pos_y")[:] = y
        env.launch_steps_random(step, 1)
        ora.fill_random_actions(step)
        ora.step()
        if step % 10 == 0 or step == 59:
            assert_same(env, ora, ctx=f"many tracks, step {step}")


def test_switching_raycast_mode_at_runtime():
    """ok_update_config may change the narrow phase of a live env: the beam tables are built on the first beam launch"""
    env, ora, tid = make_pair(["Monza", "Silverstone"], 96, 32, raycast_mode=ok.RAYCAST_GRID,
                              reward_mode=ok.REWARD_Q_PROGRESS, auto_reset=1)
    pts = spread_points(ora, tid)
    env.reset(None, pts)
    ora.reset(None, pts)
    modes = [ok.RAYCAST_GRID, ok.RAYCAST_BEAM, ok.RAYCAST_BRUTE, ok.RAYCAST_BEAM, ok.RAYCAST_GRID]
    for step in range(50):
        if step % 10 == 0:
            env.update_config(raycast_mode=modes[step // 10])
        env.launch_steps_random(step, 1)
        ora.fill_random_actions(step)
        ora.step()
        if step % 10 == 9:
            assert_same(env, ora, ctx=f"mode {modes[step // 10]}, step {step}")


@pytest.mark.parametrize("sensor_range", [120.0, 300.0])
def test_beam_mode_with_another_sensor_range(sensor_range):
    """the beam tables are built for the reference's 200 px range: a shorter range just ignores the far candidates, a
    longer one leaves the tail of every ray to the grid walk -- the results must not change"""
    env, ora, tid = make_pair(["Monza", "Catalunya"], 80, 32, raycast_mode=ok.RAYCAST_BEAM, sensor_range=sensor_range,
                              reward_mode=ok.REWARD_MIN_RAY, auto_reset=1)
    pts = spread_points(ora, tid)
    env.reset(None, pts)
    ora.reset(None, pts)
    for step in range(40):
        env.launch_steps_random(step, 1)
        ora.fill_random_actions(step)
        ora.step()
    assert_same(env, ora, ctx=f"sensor range {sensor_range}")
    assert float(ora.buffer("hit_t").max()) <= sensor_range



def test_host_step_with_pinned_buffers():
    """ok_step_host with pinned host buffers (used in place: actions read and observations written through the host
    mapping) on several tracks"""
    n = 4 * 61
    env, ora, tid = make_pair(["Monza", "Spa", "Sochi", "IMS"], n, 32, raycast_mode=ok.RAYCAST_BEAM,
                              reward_mode=ok.REWARD_CMAES_PROGRESS, auto_reset=1)
    pts = spread_points(ora, tid)
    env.reset(None, pts)
    ora.reset(None, pts)
    thr, steer = ok.pinned_array((n,), np.float32), ok.pinned_array((n,), np.float32)
    obs, rew, done = ok.pinned_array((n, 32), np.float32), ok.pinned_array((n,), np.float32), ok.pinned_array((n,), np.uint8)
    rng = np.random.default_rng(1)
    for step in range(40):
        thr[:] = rng.random(n, dtype=np.float32) * 100.0
        steer[:] = rng.random(n, dtype=np.float32) * 10.0 - 5.0
        env.step_host(thr, steer, obs, rew, done)
        ora.step(thr.copy(), steer.copy())
        assert np.array_equal(obs.view(np.uint32), ora.buffer("obs").view(np.uint32)), f"obs, step {step}"
        assert np.array_equal(rew, ora.buffer("reward")) and np.array_equal(done, ora.buffer("done")), f"step {step}"
    assert_same(env, ora, ctx="host step, pinned buffers")


@pytest.mark.parametrize("n_points", [12, 33, 40, 100])
def test_small_synthetic_tracks(n_points):
    """ovals with very few centre-line points: at most 32 points means the windowed nearest-index search never applies
    (every agent takes the full search), 33-40 points make its window wrap around most of the track"""
    from oracle.api import Oracle

    th = np.linspace(0.0, 2.0 * np.pi, n_points, endpoint=False)
    cols = (300.0 * np.cos(th), 200.0 * np.sin(th), np.full(n_points, 5.0), np.full(n_points, 5.0))
    cols = tuple(np.asarray(c, dtype=np.float32) for c in cols)
    n, rays = 48, 32
    env = ok.Env(device=0, raycast_mode=ok.RAYCAST_BEAM, reward_mode=ok.REWARD_Q_PROGRESS, auto_reset=1)
    ora = Oracle("port", reward_mode=ok.REWARD_Q_PROGRESS, auto_reset=1)
    env.add_track(cols)
    ora.add_track(cols)
    fan = ok.ray_fan(rays)
    env.alloc_agents(n, fan, None)
    ora.alloc_agents(n, fan, None)
    pts = (np.arange(n) * 5 % n_points).astype(np.int32)
    env.reset(None, pts)
    ora.reset(None, pts)
    for step in range(60):
        env.launch_steps_random(step, 1)
        ora.fill_random_actions(step)
        ora.step()
        if step % 15 == 14:
            assert_same(env, ora, ctx=f"{n_points}-point oval, step {step}")


def test_maximum_ray_count():
    """1,024 rays per agent, the most ok_alloc_agents accepts: 32 groups of rays per agent"""
    env, ora, tid = make_pair(["Zandvoort"], 6, 1024, raycast_mode=ok.RAYCAST_BEAM, reward_mode=ok.REWARD_MIN_RAY, auto_reset=1)
    pts = spread_points(ora, tid)
    env.reset(None, pts)
    ora.reset(None, pts)
    for step in range(8):
        env.launch_steps_random(step, 1)
        ora.fill_random_actions(step)
        ora.step()
    assert_same(env, ora, ctx="1024 rays")
