"""Shared helpers for the parity tests."""
import numpy as np

import openkitchen_b200 as ok
from oracle.api import Oracle

STATE_BUFS = ["pos_x", "pos_y", "rot", "speed", "accel", "act_throttle", "act_steer", "crashed", "timed_out", "done",
              "ss_ctr", "ss_x", "ss_y", "reset_pt", "start_x", "start_y"]
RAY_BUFS = ["hit_abs", "hit_rel", "obs", "hit_seg", "hit_t", "min_dist2"]
REWARD_BUFS = ["nearest_idx", "prev_idx", "reward", "fitness"]
ALL_BUFS = STATE_BUFS + RAY_BUFS + REWARD_BUFS


def spread_points(oracle, tid):
    """reset index of SURVEY 8(d): (agent_id * 2654435761 mod 2^32) mod pts"""
    return np.array([(i * 2654435761 % 2**32) % oracle.track_points(int(tid[i])) for i in range(len(tid))], dtype=np.int32)


def make_pair(track_names, n, rays_or_fan, kind="port", grouped=True, **cfg):
    """A product Env on cuda:0 and an oracle with identical tracks / agents / config."""
    fan = ok.ray_fan(rays_or_fan) if np.isscalar(rays_or_fan) else np.asarray(rays_or_fan, dtype=np.float32)
    nt = len(track_names)
    tid = (np.arange(n) * nt // n).astype(np.int32) if grouped else (np.arange(n) % nt).astype(np.int32)
    env = ok.Env(device=0, **cfg)
    ocfg = {k: v for k, v in cfg.items() if k not in ("raycast_mode", "grid_cell", "beam_cell", "beam_bins")}
    ora = Oracle(kind, **ocfg)
    for nm in track_names:
        env.add_named_track(nm)
        ora.add_track(ok.track_columns(nm))
    env.alloc_agents(n, fan, tid)
    ora.alloc_agents(n, fan, tid)
    return env, ora, tid


def same_bits(a, b):
    """elementwise: identical bits, except that any NaN equals any NaN"""
    a, b = np.asarray(a), np.asarray(b)
    if a.shape != b.shape:
        return np.zeros(1, dtype=bool)
    if a.dtype == np.float32:
        return (a.view(np.uint32) == b.view(np.uint32)) | (np.isnan(a) & np.isnan(b))
    return a == b


def assert_same(env, ora, names=ALL_BUFS, ctx=""):
    for name in names:
        a, b = env.read(name), ora.buffer(name)
        assert a.shape == b.shape, (name, a.shape, b.shape)
        if a.dtype == np.float32:
            # bit-exact, except that any NaN equals any NaN (x86 and sm_100 differ in NaN payload / sign)
            same = (a.view(np.uint32) == b.view(np.uint32)) | (np.isnan(a) & np.isnan(b))
        else:
            same = a == b
        if not same.all():
            bad = np.argwhere(~same)
            first = tuple(bad[0]) if len(bad) else None
            raise AssertionError(
                f"{ctx}: buffer {name}: {len(bad)} mismatches, first at {first}: gpu={a[first] if first else None!r} "
                f"oracle={b[first] if first else None!r}")
