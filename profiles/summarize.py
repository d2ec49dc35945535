#!/usr/bin/env python3
"""Turn ncu output brought back in gpurun_out/ into the small text summaries
kept under profiles/ (the .ncu-rep files themselves stay out of git).

  python profiles/summarize.py launches gpurun_out/launches_r1.csv > profiles/r1_launches.txt
  python profiles/summarize.py full gpurun_out/prof_r1_fused.ncu-rep > profiles/r1_search_kernel.txt
"""
import csv
import subprocess
import sys
from collections import defaultdict


def launches(path):
    rows = [r for r in csv.reader(open(path)) if len(r) > 5]
    hdr = rows[0]
    ki, mi, vi = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value")
    tot = defaultdict(lambda: [0, 0.0])
    for r in rows[1:]:
        if r[mi] != "gpu__time_duration.sum":
            continue
        name = r[ki].split("(")[0]
        tot[name][0] += 1
        tot[name][1] += float(r[vi].replace(",", ""))
    all_ns = sum(v[1] for v in tot.values())
    print("# ncu --metrics gpu__time_duration.sum --clock-control none (cold-cache, serialised: compare SHARES)")
    print("%-70s %8s %14s %8s" % ("kernel", "launches", "total_ms", "share"))
    for name, (n, ns) in sorted(tot.items(), key=lambda kv: -kv[1][1]):
        print("%-70s %8d %14.3f %7.1f%%" % (name[:70], n, ns / 1e6, 100 * ns / all_ns))


WANT = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__thread_inst_executed_per_inst_executed.ratio",
    "smsp__warps_eligible.avg.per_cycle_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "dram__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "sm__cycles_elapsed.max", "smsp__cycles_active.avg",
]


def full(path):
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    for k, vals in enumerate(rows[2:]):
        name = vals[hdr.index("Kernel Name")] if "Kernel Name" in hdr else "?"
        print(f"## launch {k}: {name}")
        for i, h in enumerate(hdr):
            if h in WANT:
                print(f"{h:70s} {vals[i]:>20s} {units[i]}")
        print("# warp stall reasons (ratio per issue-active cycle)")
        for i, h in enumerate(hdr):
            if h.startswith("smsp__average_warps_issue_stalled") and h.endswith("_per_issue_active.ratio"):
                print(f"{h:90s} {vals[i]:>10s}")
    src = subprocess.run(["ncu", "-i", path, "--page", "source", "--print-source", "cuda,sass", "--csv"],
                         capture_output=True, text=True).stdout
    cur, out = None, []
    for r in csv.reader(src.splitlines()):
        if len(r) >= 2 and r[0] in ("File Name", "File Path"):
            cur = r[1].split("/")[-1]
            continue
        if len(r) < 9 or r[0] in ("Line No", ""):
            continue
        try:
            ln, samp, inst, thr = int(r[0]), int(r[6]), int(r[7]), int(r[8])
        except ValueError:
            continue
        if inst:
            out.append((inst, samp, thr, cur, ln, r[1].strip()[:80]))
    tot = sum(x[0] for x in out)
    tots = sum(x[1] for x in out) or 1
    print("# hottest source lines: % of warp instructions, % of stall samples, active threads per instruction")
    for inst, samp, thr, f, ln, text in sorted(out, reverse=True)[:40]:
        print(f"{100 * inst / tot:5.1f}% {100 * samp / tots:5.1f}% {thr / inst:5.1f}  {f}:{ln}  {text}")


def traffic_json(path, descr, mnt):
    """{"workload", dram bytes, issue-slot utilisation, ...} of the first launch in an ncu-rep, for
    bench.py's roofline.traffic / issue.ncu (profiles/r2_sieve_kernel.json)."""
    import json
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units, vals = rows[0], rows[1], rows[2]

    def get(name):
        i = hdr.index(name)
        v = float(vals[i].replace(",", ""))
        u = units[i].lower()
        return v * (1e9 if u.startswith("gbyte") else 1e6 if u.startswith("mbyte") else 1e3 if u.startswith("kbyte") else 1)

    out = {"workload": {"descr": descr, "mnt": int(mnt)}, "kernel": vals[hdr.index("Kernel Name")],
           "dram_bytes_read": get("dram__bytes_read.sum"), "dram_bytes_write": get("dram__bytes_write.sum"),
           "gpu_time_under_ncu": vals[hdr.index("gpu__time_duration.sum")] + " " + units[hdr.index("gpu__time_duration.sum")],
           "issue": {"issue_active_pct": get("smsp__issue_active.avg.pct_of_peak_sustained_active"),
                     "alu_pipe_pct": get("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active"),
                     "warp_instructions": get("smsp__inst_executed.sum"),
                     "threads_per_instruction": get("smsp__thread_inst_executed_per_inst_executed.ratio"),
                     "warps_active_pct": get("sm__warps_active.avg.pct_of_peak_sustained_active")}}
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    if sys.argv[1] == "json":
        traffic_json(sys.argv[2], sys.argv[3], sys.argv[4])
    else:
        {"launches": launches, "full": full}[sys.argv[1]](sys.argv[2])
