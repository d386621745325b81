"""Per-tile timeline of one beam-kernel tick (ok_debug_trace): where a CTA's time goes and how the SMs finish.
usage: python tools/trace_timeline.py [agents] [cell bins]"""
import json
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
import bench  # noqa: E402
import openkitchen_b200 as ok  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
cfg = {}
if len(sys.argv) > 3:
    cfg = {"beam_cell": float(sys.argv[2]), "beam_bins": int(sys.argv[3])}
env = ok.Env(device=0, reward_mode=ok.REWARD_CMAES_PROGRESS, auto_reset=1, **cfg)
bench.build_workload(ok, env, n)
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
env.launch_steps_random(0, 100, bench.SEED)
env.sync()
res = []
for rep in range(5):
    env.debug_trace(48)
    flush.zero_()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    env.launch_steps_random(100 + rep, 1, bench.SEED, torch.cuda.current_stream().cuda_stream)
    e1.record()
    torch.cuda.synchronize()
    tr = env.debug_trace(0)
    if tr is None:
        raise SystemExit("no trace")
    used = tr[:, :, 1] > 0
    t0 = tr[:, :, 1][used].min()
    t = (tr[:, :, 1:].astype(np.int64) - int(t0)) / 1e3  # us
    start, p1, pa, pb, p4 = (t[:, :, k] for k in range(5))  # pa: thread 0's own phase 1 done (before the wait for the staged track and the barrier)
    smid = (tr[:, :, 0] >> np.uint64(32)).astype(np.int64)
    tiles_per_cta = used.sum(1)
    last_end = np.where(used, p4, 0).max(1)  # per CTA
    first_start = np.where(used, start, 1e9).min(1)
    # per-SM end
    sm_end = {}
    for c in range(tr.shape[0]):
        if used[c].any():
            sm = int(smid[c][used[c]][0])
            sm_end[sm] = max(sm_end.get(sm, 0.0), float(last_end[c]))
    ends = np.array(sorted(sm_end.values()))
    gap = np.zeros_like(start)
    gap[:, 1:] = start[:, 1:] - p4[:, :-1]
    res.append({
        "event_ms": e0.elapsed_time(e1), "span_us": float(np.where(used, p4, 0).max()), "ctas": int(tr.shape[0]),
        "tiles": int(used.sum()), "tiles_per_cta_mean": float(tiles_per_cta.mean()), "tiles_per_cta_max": int(tiles_per_cta.max()),
        "first_tile_start_us_mean": float(first_start[first_start < 1e8].mean()), "first_tile_start_us_max": float(first_start[first_start < 1e8].max()),
        "phase1_us_mean": float((p1 - start)[used].mean()), "phase1_thread0_us_mean": float((pa - start)[used].mean()),
        "rays_us_mean": float((pb - p1)[used].mean()),
        "phase4_us_mean": float((p4 - pb)[used].mean()),
        "tile_us_mean": float((p4 - start)[used].mean()),
        "between_tiles_us_mean": float(gap[:, 1:][used[:, 1:]].mean()) if used[:, 1:].any() else None,
        "sm_end_us": {"min": float(ends.min()), "p10": float(np.percentile(ends, 10)), "median": float(np.median(ends)),
                      "p90": float(np.percentile(ends, 90)), "max": float(ends.max())},
        "sm_idle_tail_frac": float(1.0 - ends.mean() / ends.max()),
    })
# per track (last repetition): tiles, agents per tile, where the time goes
tiles = env.debug_tiles()
tile_id = (tr[:, :, 0] & np.uint64(0xFFFFFFFF)).astype(np.int64)
per_track = {}
for c in range(tr.shape[0]):
    for k in range(tr.shape[1]):
        if used[c, k]:
            trk, cnt, _ = tiles[tile_id[c, k]]
            per_track.setdefault(int(trk), []).append((int(cnt), float(p1[c, k] - start[c, k]), float(pb[c, k] - p1[c, k]),
                                                       float(p4[c, k] - pb[c, k]), float(p4[c, k])))
names = ok.track_names()
res[-1]["per_track"] = {names[t]: {"tiles": len(v), "agents_per_tile": round(float(np.mean([x[0] for x in v])), 1),
                                   "phase1_us": round(float(np.mean([x[1] for x in v])), 1), "rays_us": round(float(np.mean([x[2] for x in v])), 1),
                                   "phase4_us": round(float(np.mean([x[3] for x in v])), 1),
                                   "end_us_mean": round(float(np.mean([x[4] for x in v])), 1), "end_us_max": round(float(np.max([x[4] for x in v])), 1)}
                        for t, v in sorted(per_track.items())}
print(json.dumps(res[-1], indent=1))
print(json.dumps({"event_ms_all": [r["event_ms"] for r in res], "span_us_all": [r["span_us"] for r in res]}))
