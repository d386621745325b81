"""GPU: the C++ drop-in shim and the python module `open_kitchen_pybind`.

The Template-shaped example app (openkitchen_b200/shim/example_template.cpp, written only against the
reference's class surface) is run with --trace; its action / state trace is replayed through the CPU oracle."""
import os
import subprocess

import numpy as np
import pytest

import openkitchen_b200 as ok
from oracle.api import Oracle

pytestmark = pytest.mark.gpu
LIB = os.path.join(os.path.dirname(ok.LIB_PATH))


def test_template_app_trace_replays_through_oracle(tmp_path):
    csv = tmp_path / "Monza.csv"
    ok.write_track_csv("Monza", str(csv))
    n, steps = 6, 260
    out = subprocess.run([os.path.join(LIB, "ok_shim_example"), str(csv), str(steps), str(n), "--trace"], capture_output=True,
                         text=True, timeout=120)
    assert out.returncode == 0, out.stderr
    assert "MONZA (1159 points, 4636 segments)" in out.stdout
    acts = {}
    states = {}
    for line in out.stdout.splitlines():
        f = line.split()
        if f[0] == "A":
            acts[(int(f[1]), int(f[2]))] = [np.float32(v) for v in f[3:8]] + [int(f[8])]
        elif f[0] == "S":
            states[(int(f[1]), int(f[2]))] = [np.float32(v) for v in f[3:6]] + [int(f[6]), int(f[7])] + [np.float32(v) for v in f[8:10]]
    assert len(acts) == n * steps and len(states) == n * steps
    ora = Oracle("port", movement_mode=0)
    ora.add_track(ok.track_columns("Monza"))
    ora.alloc_agents(n, np.array([-70, -30, 0, 30, 70], dtype=np.float32))
    resets = 0
    for s in range(steps):
        thr = np.array([acts[(s, a)][0] for a in range(n)], dtype=np.float32)
        st = np.array([acts[(s, a)][1] for a in range(n)], dtype=np.float32)
        # the app may have reset an agent on the host (Agent::reset) before this step: adopt the pre-step pose
        for a in range(n):
            _, _, x, y, rot, crashed = acts[(s, a)]
            if ora.buffer("crashed")[a] != crashed or ora.buffer("pos_x")[a] != x or ora.buffer("rot")[a] != rot:
                resets += 1
                ora.buffer("pos_x")[a], ora.buffer("pos_y")[a], ora.buffer("rot")[a] = x, y, rot
                ora.buffer("speed")[a] = ora.buffer("accel")[a] = 0.0
                ora.buffer("crashed")[a] = crashed
                ora.buffer("timed_out")[a] = 0
        ora.step(thr, st)
        for a in range(n):
            x, y, rot, crashed, tout, hx, hy = states[(s, a)]
            assert (ora.buffer("pos_x")[a], ora.buffer("pos_y")[a], ora.buffer("rot")[a]) == (x, y, rot), (s, a)
            assert (ora.buffer("crashed")[a], ora.buffer("timed_out")[a]) == (crashed, tout), (s, a)
            assert tuple(ora.buffer("hit_rel")[a, 2]) == (hx, hy), (s, a)
    assert resets >= n  # at least the initial resets; crashes add more
    # the stand-alone lidar with the reference's signatures (RaceTrack -> TrackSegments -> CollisionChecker(d_segments, n,
    # agents) -> checkCollision(), CollisionChecker.h:8-24): same hits and crash flags as the oracle's cast on those poses
    ora.buffer("crashed")[:] = 0
    ora.cast_rays()
    rays = [ln.split() for ln in out.stdout.splitlines() if ln.startswith("R ")]
    flags = [ln.split() for ln in out.stdout.splitlines() if ln.startswith("C ")]
    assert len(rays) == n * 5 and len(flags) == n
    hit_abs = ora.buffer("hit_abs").reshape(-1, 2)
    for f in rays:
        k = int(f[1])
        assert (np.float32(f[5]), np.float32(f[6])) == tuple(hit_abs[k]), k
        assert (np.float32(f[2]), np.float32(f[3])) == (ora.buffer("pos_x")[k // 5], ora.buffer("pos_y")[k // 5])
    for f in flags:
        a = int(f[1])
        assert int(f[2]) == int(ora.buffer("crashed")[a])
        assert (np.float32(f[3]), np.float32(f[4])) == tuple(ora.buffer("hit_rel")[a, 0])


def test_pybind_module_is_a_drop_in(tmp_path):
    import open_kitchen_pybind as okp

    csv = tmp_path / "Spa.csv"
    ok.write_track_csv("Spa", str(csv))
    env = okp.Environment(str(csv), draw_rays=False, hidden_window=True)
    env.seed(3)
    env.reset(True, False, False)
    info = env.get_render_target_info()
    assert (info.width, info.height, info.channels, info.row_bytes()) == (1600, 1400, 4, 6400)
    assert len(env.get_render_target()) == 1600 * 1400 * 4
    x0, y0, rot0, _ = env.pose()
    ora = Oracle("port", movement_mode=0)
    ora.add_track(ok.track_columns("Spa"))
    ora.alloc_agents(1, ok.ray_fan(15))  # default fan -70..70 step 10 (Agent.cpp:13-17)
    ora.buffer("pos_x")[0], ora.buffer("pos_y")[0], ora.buffer("rot")[0] = x0, y0, rot0
    for i in range(100):
        thr, st = 40.0 + (i % 7), float((i % 5) - 2)
        env.set_action(thr, st)
        env.step()
        ora.step(np.array([thr], dtype=np.float32), np.array([st], dtype=np.float32))
        if ora.buffer("crashed")[0]:
            break
    x, y, rot, speed = env.pose()
    assert (np.float32(x), np.float32(y), np.float32(rot)) == (ora.buffer("pos_x")[0], ora.buffer("pos_y")[0], ora.buffer("rot")[0])
    assert env.crashed == bool(ora.buffer("crashed")[0])
    lid = np.array(env.lidar(), dtype=np.float32)
    assert np.array_equal(lid.view(np.uint32), ora.buffer("hit_rel")[0].view(np.uint32))
    assert okp.BatchEnv is ok.BatchEnv
