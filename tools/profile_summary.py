#!/usr/bin/env python
"""Summarise ncu outputs into profiles/*.md (run in the build container; ncu reads reports without a GPU).

  python tools/profile_summary.py launches gpurun_out/launches_X.csv profiles/launches_X.md
  python tools/profile_summary.py full gpurun_out/prof_X.ncu-rep profiles/prof_X.md
"""
import csv
import io
import re
import subprocess
import sys
from collections import OrderedDict

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram__cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_tensor.sum", "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__throughput.avg.pct_of_peak_sustained_active", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_bytes.sum", "smsp__cycles_active.avg", "launch__grid_size", "launch__block_size",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio"]


def short(name):
    name = re.sub(r"unnamed>::|void |\(.*$", "", name)
    return name.replace("<", "&lt;")


def launches(src, dst):
    rows = [r for r in csv.DictReader(l for l in open(src) if l.startswith('"'))]
    per = OrderedDict()
    total = 0.0
    for r in rows:
        if r["Metric Name"] != "gpu__time_duration.sum":
            continue
        ns = float(r["Metric Value"].replace(",", ""))
        k = short(r["Kernel Name"])
        a = per.setdefault(k, [0, 0.0, r["Block Size"]])
        a[0] += 1; a[1] += ns
        total += ns
    with open(dst, "w") as f:
        f.write(f"# ncu launch list: {src}\n\n`ncu --metrics gpu__time_duration.sum --clock-control none` (cold-cache, serialised: compare SHARES)\n\n")
        f.write(f"{len(rows)} launches, {total/1e6:.3f} ms total\n\n| kernel | launches | total us | share | avg us | block |\n|---|---:|---:|---:|---:|---|\n")
        for k, (n, ns, blk) in sorted(per.items(), key=lambda kv: -kv[1][1]):
            f.write(f"| `{k}` | {n} | {ns/1e3:.1f} | {100*ns/total:.1f}% | {ns/1e3/n:.1f} | {blk} |\n")
    print(open(dst).read())


def full(src, dst):
    out = subprocess.run(["ncu", "-i", src, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    with open(dst, "w") as f:
        f.write(f"# ncu --set full: {src}\n\n")
        for r in rows[2:]:
            d = dict(zip(hdr, r))
            f.write(f"## `{short(d['Kernel Name'])}` grid {d.get('Grid Size')} block {d.get('Block Size')}\n\n| metric | value | unit |\n|---|---:|---|\n")
            for k in KEYS:
                if k in d:
                    f.write(f"| {k} | {d[k]} | {units[hdr.index(k)]} |\n")
            try:
                tr = float(d["dram__bytes_read.sum"].replace(",", "")) + float(d["dram__bytes_write.sum"].replace(",", ""))
                f.write(f"| dram traffic (read+write) | {tr:.3f} | {units[hdr.index('dram__bytes_read.sum')]} |\n")
            except (KeyError, ValueError):
                pass
            f.write("\n")
    print(open(dst).read())


if __name__ == "__main__":
    {"launches": launches, "full": full}[sys.argv[1]](sys.argv[2], sys.argv[3])
