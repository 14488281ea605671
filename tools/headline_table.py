#!/usr/bin/env python
"""Prints the headline table of profiles/README.md / DESIGN.md from profiles/rN_bench*.json (no hand-copied numbers).

  python tools/headline_table.py [r2]        (round prefix of the bench files, default r2; missing N=2 / N=8 files give empty cells)
"""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


PFX = sys.argv[1] if len(sys.argv) > 1 else "r2"


def load(name):
    path = os.path.join(ROOT, "profiles", name)
    if not os.path.exists(path):
        return {}
    with open(path) as f:
        return json.loads(f.read().strip().splitlines()[-1])


runs = [load(f"{PFX}_bench.json"), load(f"{PFX}_bench_2gpu.json"), load(f"{PFX}_bench_8gpu.json")]
f0 = lambda v: f"{v:,.0f}"


def row(label, fn):
    cells = []
    for d in runs:
        try:
            cells.append(fn(d))
        except (KeyError, TypeError):
            cells.append("")
    print(f"| {label} | " + " | ".join(cells) + " |")


print("| what | N=1 | N=2 | N=8 |\n|---|---:|---:|---:|")
row("inference B=256 per GPU, img/s (ms/step, max over ranks)", lambda d: f"{f0(d['value'])} ({d['ms_per_step']:.3f})")
row("e2e, CUDA-graph replay, fp32 NCHW in (236 MB H2D/step/GPU), uint8 mask back", lambda d: f0(d["e2e"]["value"]))
row("e2e, CUDA-graph replay, raw uint8 HWC in (59 MB H2D/step/GPU)", lambda d: f0(d["e2e"]["uint8_input"]["value"]))
row("e2e, eager `model.predict` in the same pipeline, fp32 / uint8 in",
    lambda d: f"{f0(d['e2e']['eager_predict']['value'])} / {f0(d['e2e']['uint8_input']['eager_predict']['value'])}")
row("training step, B=32 per GPU, img/s (ms/step; N>1: + 16.8 MB gradient all-reduce)",
    lambda d: f"{f0(d['train']['value'])} ({d['train']['ms_per_step']:.3f})")
row("the same step as one CUDA graph (`engine.GraphedTrainStep`, single GPU)",
    lambda d: f"{f0(d['train']['graphed']['value'])} ({d['train']['graphed']['ms_per_step']:.3f})")
row("training, global batch 256 (256 / 128 / 32 per GPU)",
    lambda d: f"{f0(d['train_global256']['value'])} ({d['train_global256']['ms_per_step']:.2f})")
row("isolated gradient all-reduce / exposed communication inside the step, ms",
    lambda d: (f"{d['train']['allreduce_isolated_ms']:.3f} / {d['train']['exposed_comm_ms']:.3f}" if d['train'].get('allreduce_isolated_ms') is not None
               else (f"{d['train']['allreduce_ms']:.3f}" if 'allreduce_ms' in d['train'] else "")))
row("global-256 training step vs the single-GPU global-256 step (`scaling_vs_n1`)",
    lambda d: f"x{d['train_global256']['scaling_vs_n1']:.2f}")
row("fp32-exact inference (B=64), img/s", lambda d: f0(d["fp32_exact"]["value"]))
row("unmodified reference on the same GPU (torch eager, cuDNN, channels_last, bf16 autocast): inference B=256 / train B=32, img/s",
    lambda d: f"{f0(d['torch_eager_gpu']['inference_b256']['value'])} / {f0(d['torch_eager_gpu']['train_step_b32']['value'])}")
row("pose head forward, B=16 per GPU, img/s (TFLOP/s per GPU, share of the measured bf16 peak)",
    lambda d: f"{f0(d['pose_head']['value'])} ({d['pose_head']['roofline']['achieved']:.0f}, {100 * d['pose_head']['roofline']['frac']:.1f} %)")
row("batch-1 latency, graph replay (configs[0] on the GPU)",
    lambda d: f"{d['latency_batch1']['ms_per_image']:.3f} ms/image" if d["n_gpus"] == 1 else "")
row("SM clock during the timed region (MHz, throttle reasons)", lambda d: f"{d['clocks']['sm_mhz']:.0f} {d['clocks']['reasons']}")
c = runs[0]["cpu_baseline"]
print(f"\nCPU baseline ({c.get('kind')}), {c['cores']} cores: {c['value']:.1f} img/s (B=32 forward), {c['config0_batch1']['value']:.1f} img/s = "
      f"{c['config0_batch1']['ms_per_image']:.2f} ms/image (B=1), training step B=32: {c['train_step_batch32']['value']:.1f} img/s "
      f"({c['train_step_batch32']['ms_per_step']:.0f} ms/step)")
r = runs[0]["roofline"]
print(f"roofline: {r['kernel']} {r['achieved']:.0f} GB/s = {100 * r['frac']:.1f} % of {r['peak']:.1f}; share of step {100 * r['share_of_step']:.1f} %; "
      f"whole step {r['whole_step_algorithmic_GB/s']:.0f} GB/s; families ms " +
      ", ".join(f"{k} {v['ms']:.3f}" for k, v in r["families"].items()))
if runs[2]:
    print(f"N=8 / N=1: x{runs[2]['value'] / runs[0]['value']:.2f}")
