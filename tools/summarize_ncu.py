#!/usr/bin/env python3
"""Summarise ncu outputs into small tracked text files under profiles/.
  launches: python tools/summarize_ncu.py launches gpurun_out/launches.csv profiles/r01_launches.md "<title>"
  full    : python tools/summarize_ncu.py full gpurun_out/prof.ncu-rep profiles/r01_full.md "<title>"
"""
import csv
import subprocess
import sys
from collections import OrderedDict


def launches(src, dst, title):
    rows = []
    with open(src) as f:
        lines = [ln for ln in f if ln.startswith('"')]
    rd = csv.reader(lines)
    hdr = next(rd)
    ix = {h: i for i, h in enumerate(hdr)}
    for r in rd:
        if r[ix["Metric Name"]] != "gpu__time_duration.sum":
            continue
        name = r[ix["Kernel Name"]].split("(")[0].replace("void ", "")
        val = float(r[ix["Metric Value"]].replace(",", ""))
        unit = r[ix["Metric Unit"]]
        ns = val * {"ns": 1, "us": 1e3, "ms": 1e6, "s": 1e9}.get(unit, 1)
        rows.append((name, ns, r[ix["Grid Size"]], r[ix["Block Size"]]))
    agg = OrderedDict()
    for name, ns, g, b in rows:
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1; a[1] += ns
    total = sum(a[1] for a in agg.values())
    with open(dst, "w") as o:
        o.write("# %s\n\nSource: `ncu --metrics gpu__time_duration.sum --clock-control none` (cold-cache, serialised launches: compare SHARES, not absolutes).\n"
                "%d launches, %.3f ms total.\n\n| kernel | launches | total ms | share |\n|---|---:|---:|---:|\n" % (title, len(rows), total / 1e6))
        for name, (n, ns) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            short = name if len(name) < 90 else name[:87] + "..."
            o.write("| `%s` | %d | %.3f | %.1f%% |\n" % (short, n, ns / 1e6, 100 * ns / total))
    print("wrote", dst)


WANT = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
        "sm__inst_executed_pipe_alu.sum", "sm__inst_executed_pipe_fma.sum", "sm__inst_executed_pipe_lsu.sum",
        "smsp__average_warp_latency_issue_stalled_long_scoreboard.ratio", "smsp__average_warp_latency_issue_stalled_short_scoreboard.ratio",
        "smsp__average_warp_latency_issue_stalled_wait.ratio", "smsp__average_warp_latency_issue_stalled_math_pipe_throttle.ratio",
        "smsp__average_warp_latency_issue_stalled_not_selected.ratio", "smsp__average_warp_latency_issue_stalled_barrier.ratio",
        "smsp__average_warp_latency_issue_stalled_lg_throttle.ratio", "smsp__average_warp_latency_issue_stalled_dispatch_stall.ratio"]


def full(src, dst, title):
    out = subprocess.run(["ncu", "-i", src, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader([ln for ln in out.splitlines() if ln.startswith('"')]))
    hdr, units = rows[0], rows[1]
    ix = {h: i for i, h in enumerate(hdr)}
    with open(dst, "w") as o:
        o.write("# %s\n\nSource: `ncu --set full --clock-control none --import-source on` (one capture per kernel; ~40 replays).\n\n" % title)
        for r in rows[2:]:
            o.write("## `%s`\n\n| metric | value | unit |\n|---|---:|---|\n" % r[ix["Kernel Name"]].split("(")[0].replace("void ", ""))
            for w in WANT:
                if w in ix:
                    o.write("| %s | %s | %s |\n" % (w, r[ix[w]], units[ix[w]]))
            o.write("\n")
    print("wrote", dst)


def traffic(src, dst, bases_per_launch, source_note):
    """profiles/r02_traffic.json for bench.py's roofline.traffic / roofline.alu: DRAM bytes per base, issue-active and thread
    instructions per window of the FIRST captured launch of each seeding kernel (bases_per_launch = bases one launch processes)."""
    import json
    out = subprocess.run(["ncu", "-i", src, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader([ln for ln in out.splitlines() if ln.startswith('"')]))
    hdr = rows[0]
    ix = {h: i for i, h in enumerate(hdr)}
    units = rows[1]

    def num(r, k):
        v = float(r[ix[k]].replace(",", ""))
        u = units[ix[k]]
        return v * {"Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "byte": 1.0}.get(u, 1.0)
    res = {}
    for r in rows[2:]:
        name = r[ix["Kernel Name"]].split("(")[0].replace("void ", "").replace("sk::", "")
        if name in res or name not in ("hashpass_kernel", "pack_kernel", "expand_kernel"):
            continue
        b = float(bases_per_launch)
        d = {"dram_bytes_per_base": (num(r, "dram__bytes_read.sum") + num(r, "dram__bytes_write.sum")) / b,
             "issue_active_pct": num(r, "smsp__issue_active.avg.pct_of_peak_sustained_active"),
             "inst_per_window": num(r, "smsp__inst_executed.sum") * num(r, "smsp__thread_inst_executed_per_inst_executed.ratio") / b,
             "bases_per_launch": b, "source": source_note}
        res[name] = d
    json.dump(res, open(dst, "w"), indent=1)
    print("wrote", dst, res)


if __name__ == "__main__":
    if sys.argv[1] == "traffic":
        traffic(sys.argv[2], sys.argv[3], sys.argv[4], sys.argv[5])
    else:
        {"launches": launches, "full": full}[sys.argv[1]](sys.argv[2], sys.argv[3], sys.argv[4])
