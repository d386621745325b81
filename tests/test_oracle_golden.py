"""CPU: the plain-C oracle (oracle/ok_oracle.c) against the golden vectors minted from the reference's
own objects (tools/make_golden.py -> tests/golden/), and -- when oracle/_ref is built -- against
those objects live.  Bit-exact on every buffer."""
import os
import sys
import zlib

import numpy as np
import pytest

import openkitchen_b200 as ok
from oracle.api import BUF, Oracle, have_ref
from tests.util import same_bits

sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "tools"))
import make_golden as mg  # noqa: E402

GOLD = os.path.join(os.path.dirname(__file__), "golden")


@pytest.fixture(scope="module")
def tracks_gold():
    return np.load(os.path.join(GOLD, "tracks.npz"))


@pytest.fixture(scope="module")
def traces_gold():
    return np.load(os.path.join(GOLD, "traces.npz"))


def test_oracle_tracks_match_golden(tracks_gold):
    o = Oracle("port")
    for nm in ok.track_names():
        t = o.add_track(ok.track_columns(nm))
        for a in mg.TRACK_ARRAYS:
            arr = o.track_array(t, a)
            assert tuple(tracks_gold[f"{nm}/{a}/shape"]) == arr.shape
            assert zlib.crc32(arr.tobytes()) == int(tracks_gold[f"{nm}/{a}/crc"][0]), (nm, a)
            assert same_bits(arr.reshape(arr.shape[0], -1)[:2], tracks_gold[f"{nm}/{a}/head"]).all()
            assert same_bits(arr.reshape(arr.shape[0], -1)[-2:], tracks_gold[f"{nm}/{a}/tail"]).all()
    assert o.track_segments(ok.track_names().index("Monza")) == 4636  # 4(P-1)+4
    assert sum(o.track_segments(i) for i in range(23)) == 90676


@pytest.mark.parametrize("name", list(mg.SCENARIOS))
def test_oracle_traces_match_golden(name, traces_gold):
    got = mg.run_scenario("port", name)
    assert got, "scenario produced no checkpoints"
    for key, arr in got.items():
        assert same_bits(arr, traces_gold[key]).all(), key


@pytest.mark.skipif(not have_ref(), reason="oracle/_ref not built (needs /root/reference)")
@pytest.mark.parametrize("mode,reward", [(0, 1), (0, 2), (1, 7), (1, 5), (0, 4), (0, 3), (1, 6)])
def test_oracle_vs_reference_objects_live(mode, reward):
    names = ["Monza", "Zandvoort", "Spa"]
    pair = []
    for kind in ("port", "reference"):
        o = Oracle(kind, movement_mode=mode, reward_mode=reward, auto_reset=1)
        for nm in names:
            o.add_track(ok.track_columns(nm))
        tid = (np.arange(60) % 3).astype(np.int32)
        o.alloc_agents(60, ok.ray_fan(15), tid)
        pair.append(o)
    a, b = pair
    rng = np.random.default_rng(7)
    idx = np.arange(60, dtype=np.int64)
    pts = np.array([rng.integers(0, a.track_points(int(i % 3))) for i in idx], dtype=np.int32)
    alpha = (rng.integers(10, 91, size=60) / np.float32(100)).astype(np.float32)
    hoff = rng.integers(-90, 91, size=60).astype(np.float32)
    for o in pair:
        o.reset(idx, pts, alpha, hoff)
    for step in range(150):
        for o in pair:
            o.fill_random_actions(step)
            o.step()
        if step % 25 == 0 or step == 149:
            for name in BUF:
                assert same_bits(a.buffer(name), b.buffer(name)).all(), (step, name)


@pytest.mark.skipif(not os.path.isdir("/root/reference/tracks"), reason="reference CSVs not present")
def test_track_pack_matches_reference_csvs(tmp_path):
    """the packed columns are exactly what strtof parses out of the reference's CSV files, and the
    CSV writer round-trips them"""
    o = Oracle("port")
    for nm in ok.track_names():
        t_csv = o.load_track_csv(f"/root/reference/tracks/{nm}.csv")
        t_pack = o.add_track(ok.track_columns(nm))
        p = tmp_path / f"{nm}.csv"
        ok.write_track_csv(nm, str(p))
        t_rt = o.load_track_csv(str(p))
        for a in ("x", "y", "w_right", "w_left", "segments"):
            ref = o.track_array(t_csv, a)
            assert same_bits(ref, o.track_array(t_pack, a)).all(), (nm, a)
            assert same_bits(ref, o.track_array(t_rt, a)).all(), (nm, a)


@pytest.mark.skipif(not have_ref(), reason="oracle/_ref not built (needs /root/reference)")
@pytest.mark.parametrize("mode", [0, 1])
@pytest.mark.parametrize("offset", [3.0, -5.0])
def test_oracle_vs_reference_objects_with_a_sensor_offset(mode, offset):
    """Agent::sensor_offset_ (Agent.h:69, CollisionChecker.cu:122-126: the lidar origin is shifted along the heading) and the
    crash threshold of CollisionChecker.cu:167 -- what the reference's objects let one vary (dt, speed limit, sensor range
    and the standstill window are compile-time constants there: Agent.h:10-12, Environment.h:19-20) -- in the C restatement
    against the reference's own objects.  The GPU tests compare the CUDA path with the restatement on the same settings
    (tests/test_gpu_parity_gaps.py): this closes the chain CUDA == restatement == reference for them."""
    names = ["Monza", "Sepang"]
    cfg = dict(movement_mode=mode, reward_mode=1, auto_reset=1, sensor_offset=offset, collision_dist2=3.5)
    pair = []
    for kind in ("port", "reference"):
        o = Oracle(kind, **cfg)
        for nm in names:
            o.add_track(ok.track_columns(nm))
        o.alloc_agents(48, ok.ray_fan(9), (np.arange(48) % 2).astype(np.int32))
        pair.append(o)
    a, b = pair
    rng = np.random.default_rng(11)
    idx = np.arange(48, dtype=np.int64)
    pts = np.array([rng.integers(0, a.track_points(int(i % 2))) for i in idx], dtype=np.int32)
    for o in pair:
        o.reset(idx, pts, None, None)
    crashed_seen = False
    ticks = 120 if mode == 0 else 400  # the acceleration mode's action set builds speed slowly
    for step in range(ticks):
        for o in pair:
            o.fill_random_actions(step)
            o.step()
        if step % 20 == 0 or step == ticks - 1:
            for name in BUF:
                assert same_bits(a.buffer(name), b.buffer(name)).all(), (step, name)
        crashed_seen = crashed_seen or bool(a.buffer("crashed").any())
    assert crashed_seen, "nobody crashed: the changed threshold was never exercised"
