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


def fullcsv(src, dst):
    """Same table from a CSV that was exported on the GPU box (`ncu -i X.ncu-rep --page raw --csv > X.csv`; the report itself is too
    large to bring back), one compact row per launch."""
    rows = list(csv.reader(l for l in open(src) if l.startswith('"')))
    hdr, units = rows[0], rows[1]
    col = {n: i for i, n in enumerate(hdr)}
    sel = [("gpu__time_duration.sum", "us", 1e-3), ("dram__bytes_read.sum", "rd MB", None), ("dram__bytes_write.sum", "wr MB", None),
           ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram thr %", 1), ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm thr %", 1),
           ("l1tex__throughput.avg.pct_of_peak_sustained_active", "l1tex %", 1), ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "L2 thr %", 1),
           ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps act %", 1), ("launch__registers_per_thread", "regs", 1),
           ("launch__occupancy_limit_registers", "occ regs", 1), ("launch__occupancy_limit_shared_mem", "occ smem", 1)]

    def num(r, k):
        try:
            return float(r[col[k]].replace(",", ""))
        except (KeyError, ValueError):
            return float("nan")

    def mb(r, k):  # ncu scales byte columns per column (unit row): normalise to MB
        u = units[col[k]].lower() if k in col else ""
        f = {"byte": 1e-6, "kbyte": 1e-3, "mbyte": 1.0, "gbyte": 1e3}.get(u, 1e-6)
        return num(r, k) * f
    with open(dst, "w") as f:
        f.write(f"# ncu --set full (CSV export): {src}\n\n`--clock-control none`; caches are flushed before every kernel, launches are serialised.\n\n")
        f.write("| kernel | grid | " + " | ".join(n for _, n, _ in sel) + " | GB/s (rd+wr) |\n|---|---|" + "---:|" * (len(sel) + 1) + "\n")
        for r in rows[2:]:
            us = num(r, "gpu__time_duration.sum") * ({"ns": 1e-3, "nsecond": 1e-3, "us": 1.0, "usecond": 1.0, "ms": 1e3, "msecond": 1e3}.get(units[col["gpu__time_duration.sum"]].lower(), 1e-3))
            vals = []
            for k, _, sc in sel:
                if k == "gpu__time_duration.sum":
                    vals.append(f"{us:.1f}")
                elif sc is None:
                    vals.append(f"{mb(r, k):.1f}")
                else:
                    vals.append(f"{num(r, k):.1f}")
            gbs = (mb(r, "dram__bytes_read.sum") + mb(r, "dram__bytes_write.sum")) / us * 1e3 if us > 0 else float("nan")  # MB/us = TB/s
            f.write(f"| `{short(r[col['Kernel Name']])[:60]}` | {r[col['Grid Size']] if 'Grid Size' in col else ''} | " + " | ".join(vals) + f" | {gbs:.0f} |\n")
    print(open(dst).read())


if __name__ == "__main__":
    {"launches": launches, "full": full, "fullcsv": fullcsv}[sys.argv[1]](sys.argv[2], sys.argv[3])
