#!/usr/bin/env python
"""profiles/traffic.json (what bench.py's `roofline` cites) from one `ncu --set full` capture of the step kernel and the
launch list of the same command.  usage: python tools/make_traffic.py <prof.ncu-rep> <launches.csv> <summary name> [out.json ...]"""
import collections
import csv
import io
import json
import subprocess
import sys

rep, launches, summary = sys.argv[1:4]
outs = sys.argv[4:] or ["profiles/traffic.json"]
raw = list(csv.reader(io.StringIO(subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout)))
hdr, row = raw[0], raw[2]
get = lambda k: float(row[hdr.index(k)].replace(",", ""))  # noqa: E731
kernel = row[hdr.index("Kernel Name")]
share, count = collections.Counter(), collections.Counter()
with open(launches) as f:
    lines = [ln for ln in f if ln.startswith('"')]
for r in csv.DictReader(lines):
    if r.get("Metric Name") == "gpu__time_duration.sum":
        share[r["Kernel Name"]] += float(r["Metric Value"].replace(",", ""))
        count[r["Kernel Name"]] += 1
rd, wr = get("dram__bytes_read.sum"), get("dram__bytes_write.sum")
unit = raw[1][hdr.index("dram__bytes_read.sum")]
scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
rd *= scale[unit]
wr *= scale[raw[1][hdr.index("dram__bytes_write.sum")]]
dur = get("gpu__time_duration.sum") * {"ns": 1e-3, "us": 1.0, "ms": 1e3}[raw[1][hdr.index("gpu__time_duration.sum")]]
doc = {
    "kernel": f"{kernel} (beam mode, staged shape, feedback tiling)",
    "dram_bytes_per_launch": rd + wr, "dram_bytes_read": rd, "dram_bytes_write": wr,
    "warp_instructions_per_launch": int(get("smsp__inst_executed.sum")),
    "issue_active_pct": round(get("smsp__issue_active.avg.pct_of_peak_sustained_active"), 2),
    "active_threads_per_instruction": round(get("smsp__thread_inst_executed_per_inst_executed.ratio"), 2),
    "duration_us_under_ncu": dur,
    "source": f"profiles/{summary} (ncu --set full, one launch of `python bench.py`, 65,536 agents x 32 rays, beam table 2 px x 256 bins, round 2)",
    "launch_list": "profiles/r2_launches.csv",
    "time_share_ns": dict(share), "launches": dict(count),
}
# the sentence bench.py prints next to the roofline numbers, from this capture's own metrics
stalls = {}
try:
    for ln in open(f"profiles/{summary}"):
        if ln.startswith("stall_"):
            k, v = ln.split()
            stalls[k] = float(v.rstrip("%"))
except OSError:
    pass
top = sorted(stalls.items(), key=lambda kv: -kv[1])[:3]
pipe = lambda k: round(get(f"sm__inst_executed_pipe_{k}.avg.pct_of_peak_sustained_active"))  # noqa: E731
doc["note"] = (f"not HBM bound: warp-issue bound (ncu: {doc['issue_active_pct']:.1f} % of issue slots busy while a scheduler has work, "
               f"{doc['active_threads_per_instruction']} active threads per instruction, ALU pipe ~{pipe('alu')} %, FMA ~{pipe('fma')} %, "
               f"XU ~{pipe('xu')} %, FP64 ~{pipe('fp64')} %, top stalls " + ", ".join(f"{k[6:]} {v:.1f} %" for k, v in top) +
               f"); DRAM traffic ({(rd + wr) / 1e6:.0f} MB) exceeds the algorithmic 22.8 MB because every ray fetches one 16-byte table entry as a "
               f"random sector of a 5.7 GB table -- profiles/{summary}, profiles/r2_step_kernel_phases.txt, profiles/r2_timeline.json")
for o in outs:
    json.dump(doc, open(o, "w"), indent=1)
print(json.dumps(doc, indent=1))
