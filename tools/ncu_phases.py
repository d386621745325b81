#!/usr/bin/env python
"""Split an .ncu-rep's SASS profile of the step kernel at its CTA barriers (BAR.SYNC): instructions, samples and
lane occupancy per phase.  usage: python tools/ncu_phases.py gpurun_out/prof.ncu-rep"""
import csv
import io
import subprocess
import sys


def main():
    rep = sys.argv[1]
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr = None
    segs, cur = [], {"n": 0, "inst": 0, "thr": 0, "samp": 0, "first": None, "marks": []}
    for r in rows:
        if len(r) > 5 and r[0] == "Address":
            hdr = r
            continue
        if hdr is None or len(r) < len(hdr) - 2:
            continue
        d = dict(zip(hdr, r))
        try:
            inst, thr, samp = int(d["Instructions Executed"]), int(d["Thread Instructions Executed"]), int(d["# Samples"])
        except Exception:
            continue
        src = d["Source"].strip()
        cur["n"] += 1
        cur["inst"] += inst
        cur["thr"] += thr
        cur["samp"] += samp
        for key in ("SHFL", "LDG", "ATOMS", "ATOMG", "UBLKCP", "MUFU", "DFMA", "SYNCS"):
            if key in src and key not in cur["marks"]:
                cur["marks"].append(key)
        if "BAR.SYNC" in src or "EXIT" in src:
            cur["end"] = src[:40]
            segs.append(cur)
            cur = {"n": 0, "inst": 0, "thr": 0, "samp": 0, "first": None, "marks": []}
    if cur["n"]:
        cur["end"] = "(end)"
        segs.append(cur)
    ti = sum(s["inst"] for s in segs) or 1
    ts = sum(s["samp"] for s in segs) or 1
    print(f"total warp instructions {ti}, samples {ts}")
    for i, s in enumerate(segs):
        if s["inst"] == 0 and s["samp"] == 0:
            continue
        print(f"seg {i:2d}: sass {s['n']:5d}  inst {100 * s['inst'] / ti:5.1f}%  samples {100 * s['samp'] / ts:5.1f}%  lanes {s['thr'] / max(s['inst'], 1):5.1f}  "
              f"ends {s['end']:<28s} has {','.join(s['marks'])}")


if __name__ == "__main__":
    main()
