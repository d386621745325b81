#!/usr/bin/env python
"""Summarise an .ncu-rep: headline metrics, stall reasons, opcode mix and the hottest source lines.
usage: python tools/ncu_summary.py gpurun_out/prof.ncu-rep [out_prefix]"""
import collections
import csv
import io
import subprocess
import sys


def ncu(rep, *args):
    return subprocess.run(["ncu", "-i", rep, *args], capture_output=True, text=True).stdout


def main():
    rep = sys.argv[1]
    out = []
    raw = list(csv.reader(io.StringIO(ncu(rep, "--page", "raw", "--csv"))))
    hdr, units = raw[0], raw[1]
    want = ["gpu__time_duration.sum", "sm__cycles_elapsed.max", "smsp__cycles_active.avg", "smsp__inst_executed.sum",
            "smsp__thread_inst_executed_per_inst_executed.ratio", "smsp__issue_active.avg.pct_of_peak_sustained_active",
            "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__warps_eligible.avg.per_cycle_active",
            "launch__registers_per_thread", "launch__block_size", "launch__grid_size",
            "launch__shared_mem_per_block_dynamic", "dram__bytes_read.sum", "dram__bytes_write.sum",
            "dram__throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
            "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
            "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
            "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
            "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
            "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
            "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active"]
    out.append("== headline metrics (per launch) ==")
    for k in want:
        if k in hdr:
            i = hdr.index(k)
            out.append(f"{k:75s} {units[i]:14s} " + " ".join(r[i] for r in raw[2:]))
    cs = list(csv.reader(io.StringIO(ncu(rep, "--page", "source", "--csv", "--print-source", "cuda,sass"))))
    hdr2, cur = None, None
    agg = collections.OrderedDict()
    stalls = collections.Counter()
    ops = collections.Counter()
    for r in cs:
        if len(r) == 2 and r[0] == "File Path":
            cur = r[1].split("/")[-1]
            continue
        if len(r) > 10 and r[0] == "Line No":
            hdr2 = r
            continue
        if hdr2 is None or len(r) < 10:
            continue
        d = dict(zip(hdr2, r))
        try:
            inst, thr, samp = int(d["Instructions Executed"]), int(d["Thread Instructions Executed"]), int(d["# Samples"])
        except Exception:
            continue
        if r[2] == "-":
            a = agg.setdefault((cur, int(r[0]), r[1].strip()[:84]), [0, 0, 0])
            a[0] += inst
            a[1] += thr
            a[2] += samp
        else:
            op = r[3].split()
            if op:
                o = op[1] if op[0].startswith("@") and len(op) > 1 else op[0]
                ops[o.split(".")[0]] += inst
            for h in hdr2:
                if h.startswith("stall_") and "Not Issued" not in h:
                    try:
                        stalls[h] += int(d[h])
                    except Exception:
                        pass
    tot = sum(a[0] for a in agg.values()) or 1
    tots = sum(a[2] for a in agg.values()) or 1
    thr_tot = sum(a[1] for a in agg.values())
    out.append(f"\n== totals: warp instructions {tot}, avg active threads {thr_tot / tot:.2f}, samples {tots} ==")
    out.append("\n== stall reasons (share of samples) ==")
    st = sum(stalls.values()) or 1
    for h, v in stalls.most_common(10):
        out.append(f"{h:28s} {100 * v / st:5.1f}%")
    out.append("\n== opcode mix (share of warp instructions) ==")
    ot = sum(ops.values()) or 1
    for o, v in ops.most_common(18):
        out.append(f"{o:10s} {100 * v / ot:5.1f}%")
    out.append("\n== hottest source lines ==")
    for k, a in sorted(agg.items(), key=lambda kv: -kv[1][0])[:40]:
        out.append(f"{k[0][:15]:15s} L{k[1]:4d} inst {100 * a[0] / tot:5.1f}% samp {100 * a[2] / tots:5.1f}% thr/inst {a[1] / max(a[0], 1):5.1f} | {k[2]}")
    text = "\n".join(out)
    print(text)
    if len(sys.argv) > 2:
        open(sys.argv[2], "w").write(text + "\n")


if __name__ == "__main__":
    main()
