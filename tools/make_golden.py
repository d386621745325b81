#!/usr/bin/env python
"""Mint the golden vectors under tests/golden/ from the REFERENCE ITSELF run here:
oracle/_ref/libokref.so = the unmodified reference Agent.cpp + RaceTrack.cpp driven by
oracle/ref_harness.cpp (see oracle/Makefile).  The reference ships no tests or fixtures of its own
(SURVEY.md section 4), so these are the pins for the oracle and, through it, for the CUDA path.

    python tools/make_golden.py            # needs /root/reference (build with `make -C oracle`)

Contents
  tracks.npz   per track: CRC32 of every RaceTrack / TrackSegments array + first/last rows
  traces.npz   per scenario: full state + lidar buffers at a few checkpoints of a scripted rollout
"""
import os
import sys
import zlib

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import openkitchen_b200 as ok  # noqa: E402  (track pack + ray fan helpers only)
from oracle.api import BUF, Oracle  # noqa: E402

TRACK_ARRAYS = ["x", "y", "w_right", "w_left", "heading", "li", "lo", "ri", "ro", "segments"]

# name: (tracks, n_agents, rays/fan, config, ticks, checkpoints)
SCENARIOS = {
    # BASELINE config 0: one agent on Monza, Q-learning fan, VELOCITY mode
    "c1_single_monza": (["Monza"], 1, [-70, -30, 0, 30, 70], dict(movement_mode=0, reward_mode=1, auto_reset=1), 1000,
                        [0, 1, 10, 100, 500, 999]),
    # BASELINE config 1 shape (reduced population): Silverstone, 32 rays, ACCELERATION mode
    "c2_silverstone_accel": (["Silverstone"], 64, 32, dict(movement_mode=1, reward_mode=6, auto_reset=0), 400,
                             [0, 50, 200, 201, 399]),
    # BASELINE config 2 shape (reduced population): all 23 tracks, 32 rays, random actions, auto-reset
    "c3_all_tracks": (None, 92, 32, dict(movement_mode=0, reward_mode=2, auto_reset=1), 300, [0, 20, 100, 299]),
    # default 15-ray fan, lane-centre reward, no auto reset (agents crash and keep stale hits)
    "default_fan_lane_center": (["Spa", "Austin"], 40, 15, dict(movement_mode=0, reward_mode=7, auto_reset=0), 250,
                                [0, 30, 249]),
}


def fan_of(spec):
    return ok.ray_fan(spec) if np.isscalar(spec) else np.asarray(spec, dtype=np.float32)


def run_scenario(kind, name):
    tracks, n, rays, cfg, ticks, cps = SCENARIOS[name]
    tracks = tracks or ok.track_names()
    o = Oracle(kind, **cfg)
    for t in tracks:
        o.add_track(ok.track_columns(t))
    tid = (np.arange(n) * len(tracks) // n).astype(np.int32)
    o.alloc_agents(n, fan_of(rays), tid)
    pts = np.array([(i * 2654435761 % 2**32) % o.track_points(int(tid[i])) for i in range(n)], dtype=np.int32)
    o.reset(None, pts)
    out = {}
    for s in range(ticks):
        o.fill_random_actions(s)
        o.step()
        if s in cps:
            for b in BUF:
                out[f"{name}/{s}/{b}"] = o.buffer(b).copy()
    return out


def main():
    ref = Oracle("reference")
    tr = {}
    for nm in ok.track_names():
        t = ref.add_track(ok.track_columns(nm))
        for a in TRACK_ARRAYS:
            arr = ref.track_array(t, a)
            tr[f"{nm}/{a}/crc"] = np.array([zlib.crc32(arr.tobytes())], dtype=np.uint32)
            tr[f"{nm}/{a}/shape"] = np.array(arr.shape, dtype=np.int32)
            tr[f"{nm}/{a}/head"] = arr.reshape(arr.shape[0], -1)[:2].copy()
            tr[f"{nm}/{a}/tail"] = arr.reshape(arr.shape[0], -1)[-2:].copy()
    os.makedirs(os.path.join(ROOT, "tests", "golden"), exist_ok=True)
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "tracks.npz"), **tr)
    traces = {}
    for name in SCENARIOS:
        traces.update(run_scenario("reference", name))
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "traces.npz"), **traces)
    for f in ("tracks.npz", "traces.npz"):
        print(f, os.path.getsize(os.path.join(ROOT, "tests", "golden", f)), "bytes")


if __name__ == "__main__":
    main()
