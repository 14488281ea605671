#!/usr/bin/env python
"""Per-layer summary of an `ncu --set full` capture of ONE forward (conv kernels only), joined with the live per-layer
event times of the same tree.

  ncu --set full --clock-control none -k regex:"dwconv_smem_kernel|conv_gemm_kernel" -c 46 -o /tmp/p \
      python bench.py --no-graph --steps 1 --warmup 1 --no-cpu --no-train --no-pose
  ncu -i /tmp/p.ncu-rep --page raw --csv > gpurun_out/prof_fwd_X.csv
  python tools/forward_full_summary.py gpurun_out/prof_fwd_X.csv gpurun_out/layers_X.json profiles/rN_forward_full.md profiles/rN_traffic.json

The first 46 matching launches of the process are one forward in layer order (every forward is identical); the join is
checked kernel-kind by kernel-kind and aborts on a mismatch.
"""
import csv
import json
import re
import sys


def kind_of(name):
    if "dwconv" in name or "dw_col_kernel" in name:
        return "dwconv"
    if "conv3x3_halo_kernel" in name:
        return "conv_gemm_3x3"
    m = re.search(r"conv_gemm_kernel<\s*(\d)\s*,\s*(\d)\s*>", name)
    if not m:
        return "?"
    if m.group(1) == "1":
        return "conv_gemm_3x3"
    return "conv_gemm_1x1_se" if m.group(2) == "1" else "conv_gemm_1x1"


def main(src, layers_json, dst_md, dst_json):
    rows = list(csv.reader(l for l in open(src) if l.startswith('"')))
    head, units, kernels = rows[0], rows[1], rows[2:]
    col = {n: i for i, n in enumerate(head)}

    def val(r, n):
        v = r[col[n]].replace(",", "")
        try:
            return float(v)
        except ValueError:
            return float("nan")

    def mbytes(r, n):  # ncu picks a unit per column
        u = units[col[n]].lower()
        scale = {"byte": 1e-6, "kbyte": 1e-3, "mbyte": 1.0, "gbyte": 1e3}[u]
        return val(r, n) * scale

    assert units[col["gpu__time_duration.sum"]] in ("us", "usecond"), units[col["gpu__time_duration.sum"]]
    layers = [e for e in json.load(open(layers_json))["layers"] if e["kernel"].startswith(("conv_gemm", "dwconv"))]
    assert len(layers) == len(kernels), (len(layers), len(kernels))
    out, fam = [], {}
    for r, e in zip(kernels, layers):
        name = r[col["Kernel Name"]]
        assert kind_of(name) == e["kernel"], (name, e)
        rd, wr = mbytes(r, "dram__bytes_read.sum"), mbytes(r, "dram__bytes_write.sum")
        rec = {
            "layer": e["name"], "kernel": e["kernel"], "template": (re.search(r"(\w+_kernel<[^>]*>)", name) or re.search(r"(\w+_kernel)", name)).group(1),
            "grid": r[col["Grid Size"]], "block": r[col["Block Size"]],
            "ncu_us": val(r, "gpu__time_duration.sum"), "event_us": e["ms"] * 1e3,
            "dram_read_MB": rd, "dram_write_MB": wr, "traffic_MB": rd + wr, "algorithmic_MB": e["bytes"] / 1e6,
            "warps_active_pct": val(r, "sm__warps_active.avg.pct_of_peak_sustained_active"),
            "sm_throughput_pct": val(r, "sm__throughput.avg.pct_of_peak_sustained_elapsed"),
            "tensor_pipe_pct": val(r, "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"),
            "lts_throughput_pct": val(r, "lts__throughput.avg.pct_of_peak_sustained_elapsed"),
            "dram_throughput_pct": val(r, "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"),
            "regs": val(r, "launch__registers_per_thread"),
            "limit_regs": val(r, "launch__occupancy_limit_registers"),
            "limit_smem": val(r, "launch__occupancy_limit_shared_mem"),
        }
        out.append(rec)
        f = fam.setdefault(e["kernel"], {"launches": 0, "traffic_MB": 0.0, "algorithmic_MB": 0.0, "ncu_us": 0.0, "event_us": 0.0})
        f["launches"] += 1
        for k in ("traffic_MB", "algorithmic_MB", "ncu_us", "event_us"):
            f[k] += rec[k]
    for f in fam.values():
        f["traffic_bytes_per_launch"] = f["traffic_MB"] * 1e6 / f["launches"]
        f["traffic_over_algorithmic"] = f["traffic_MB"] / f["algorithmic_MB"]
    json.dump({"source": src, "layers_source": layers_json, "note": "dram__bytes_read.sum + dram__bytes_write.sum per launch, ncu --set full, "
               "cache flushed before every kernel (ncu default): bytes still dirty in L2 when a kernel ends are NOT counted",
               "families": fam, "layers": out}, open(dst_json, "w"), indent=1)
    with open(dst_md, "w") as fh:
        fh.write(f"# ncu --set full, one B=256 forward, conv kernels in layer order: {src}\n\n"
                 f"Joined with the live CUDA-event layer times of the same tree ({layers_json}). ncu flushes the caches before every\n"
                 "kernel and serialises launches, so `ncu us` is a cold, isolated launch; `event us` is the launch inside the running forward.\n"
                 "`traffic` = dram read + write during the kernel; output bytes still dirty in L2 at kernel end are not in it.\n\n"
                 "| layer | kernel | grid | ncu us | event us | dram rd MB | dram wr MB | traffic MB | algorithmic MB | traffic/alg | warps act % | sm thr % | tensor % | L2 thr % | dram thr % | regs | occ limit regs/smem |\n"
                 "|---|---|---|---:|---:|---:|---:|---:|---:|---:|---:|---:|---:|---:|---:|---:|---|\n")
        for r in out:
            fh.write(f"| {r['layer']} | `{r['template'].replace('<', '&lt;')}` | {r['grid']} | {r['ncu_us']:.1f} | {r['event_us']:.1f} | {r['dram_read_MB']:.1f} | "
                     f"{r['dram_write_MB']:.1f} | {r['traffic_MB']:.1f} | {r['algorithmic_MB']:.1f} | {r['traffic_MB'] / r['algorithmic_MB']:.2f} | "
                     f"{r['warps_active_pct']:.1f} | {r['sm_throughput_pct']:.1f} | {r['tensor_pipe_pct']:.1f} | {r['lts_throughput_pct']:.1f} | "
                     f"{r['dram_throughput_pct']:.1f} | {r['regs']:.0f} | {r['limit_regs']:.0f}/{r['limit_smem']:.0f} |\n")
        fh.write("\n## Families\n\n| family | launches | traffic MB | algorithmic MB | traffic/alg | ncu us | event us |\n|---|---:|---:|---:|---:|---:|---:|\n")
        for k, f in fam.items():
            fh.write(f"| {k} | {f['launches']} | {f['traffic_MB']:.1f} | {f['algorithmic_MB']:.1f} | {f['traffic_over_algorithmic']:.2f} | {f['ncu_us']:.1f} | {f['event_us']:.1f} |\n")


if __name__ == "__main__":
    main(*sys.argv[1:5])
