#!/usr/bin/env python
"""Times the 15 depthwise layers of the B=256 forward in isolation through the C ABI (CUDA events, rotating input
buffers so that no launch finds its input in L2), and checks every variant against the default one.

  MTGSEG_DW_VARIANT=<v> python tools/dw_probe.py [out.json]      (the variant is read once per process)
"""
import json, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import devops as D
from mtg_card_image_segmentation_b200 import _native as N

B = int(os.environ.get("DW_B", "256"))
REPS = 6
LAYERS = [("b1", 160, 120, 16, 3, 1, 1, 1, 0), ("b2", 160, 120, 64, 3, 2, 1, 1, 0), ("b3", 80, 60, 72, 3, 1, 1, 1, 0),
          ("b4", 80, 60, 72, 5, 2, 1, 1, 1), ("b5", 40, 30, 120, 5, 1, 1, 1, 1), ("b6", 40, 30, 120, 5, 1, 1, 1, 1),
          ("b7", 40, 30, 240, 3, 2, 1, 2, 0), ("b8", 20, 15, 200, 3, 1, 1, 2, 0), ("b9", 20, 15, 184, 3, 1, 1, 2, 0),
          ("b10", 20, 15, 184, 3, 1, 1, 2, 0), ("b11", 20, 15, 480, 3, 1, 1, 2, 1), ("b12", 20, 15, 672, 3, 1, 1, 2, 1),
          ("b13", 20, 15, 672, 5, 1, 2, 2, 1), ("b14", 20, 15, 960, 5, 1, 2, 2, 1), ("b15", 20, 15, 960, 5, 1, 2, 2, 1)]
ONLY = [x for x in os.environ.get("DW_LAYERS", "").split(",") if x]
if ONLY:
    LAYERS = [l for l in LAYERS if l[0] in ONLY]
REPS = int(os.environ.get("DW_REPS", REPS))
dev = "cuda"
variant = os.environ.get("MTGSEG_DW_VARIANT", "0")
res = {"variant": variant, "batch": B, "layers": {}}
total = 0.0
g = torch.Generator(device=dev).manual_seed(0)
for (name, H, W, C, k, s, d, act, gap) in LAYERS:
    nbuf = 3
    xs = [torch.randn(B, H, W, C, device=dev, generator=g).bfloat16() for _ in range(nbuf)]
    w = (torch.randn(k * k, C, device=dev, generator=g) * 0.2).bfloat16()
    sc = torch.rand(C, device=dev, generator=g) + 0.5
    sh = torch.randn(C, device=dev, generator=g) * 0.1
    if os.environ.get("DW_NCU"):  # one launch per layer for a profiler capture
        D.dwconv(xs[0], w, sc, sh, act, k, s, d, bool(gap)); torch.cuda.synchronize()
        del xs; torch.cuda.empty_cache()
        continue
    for i in range(nbuf):
        out, gp = D.dwconv(xs[i], w, sc, sh, act, k, s, d, bool(gap))
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for r in range(REPS):
        D.dwconv(xs[r % nbuf], w, sc, sh, act, k, s, d, bool(gap))
    e1.record(); torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 1e3 / REPS
    # checksum of the last buffer's output for cross-variant comparison (bit-exact expected: same arithmetic order)
    out, gp = D.dwconv(xs[0], w, sc, sh, act, k, s, d, bool(gap))
    chk = float(out.float().abs().sum().item()); gchk = float(gp.sum().item()) if gp is not None else 0.0
    algo = (xs[0].numel() + out.numel()) * 2
    res["layers"][name] = {"us": us, "GB/s": algo / us / 1e3, "abs_sum": chk, "gap_sum": gchk}
    total += us
    del xs, out
    torch.cuda.empty_cache()
res["total_us"] = total
print(f"variant {variant}: total {total:.1f} us | " + " ".join(f"{k}={v['us']:.0f}" for k, v in res["layers"].items()))
if len(sys.argv) > 1:
    json.dump(res, open(sys.argv[1], "w"), indent=1)
