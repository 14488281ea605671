#!/usr/bin/env python
"""Times the 31 pointwise convs of the B=256 forward in isolation through the C ABI (CUDA events, rotating input buffers
so that no launch finds its input in L2).  MTGSEG_GEMM_DEBUG=1 prints the launch configuration of every call.

  python tools/gemm_probe.py [out.json]
"""
import json, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import devops as D

B = int(os.environ.get("GEMM_B", "256"))
REPS = 6
# name, pixels per image, K, N, act (0 none / 1 relu / 2 hardswish), residual, squeeze-excite gate on the input
L = [("b1.project", 160 * 120, 16, 16, 0, 1, 0), ("b2.expand", 160 * 120, 16, 64, 1, 0, 0), ("b2.project", 80 * 60, 64, 24, 0, 0, 0),
     ("b3.expand", 80 * 60, 24, 72, 1, 0, 0), ("b3.project", 80 * 60, 72, 24, 0, 1, 0), ("b4.expand", 80 * 60, 24, 72, 1, 0, 0),
     ("b4.project", 40 * 30, 72, 40, 0, 0, 1), ("b5.expand", 40 * 30, 40, 120, 1, 0, 0), ("b5.project", 40 * 30, 120, 40, 0, 1, 1),
     ("b6.expand", 40 * 30, 40, 120, 1, 0, 0), ("b6.project", 40 * 30, 120, 40, 0, 1, 1), ("b7.expand", 40 * 30, 40, 240, 2, 0, 0),
     ("b7.project", 20 * 15, 240, 80, 0, 0, 0), ("b8.expand", 20 * 15, 80, 200, 2, 0, 0), ("b8.project", 20 * 15, 200, 80, 0, 1, 0),
     ("b9.expand", 20 * 15, 80, 184, 2, 0, 0), ("b9.project", 20 * 15, 184, 80, 0, 1, 0), ("b10.expand", 20 * 15, 80, 184, 2, 0, 0),
     ("b10.project", 20 * 15, 184, 80, 0, 1, 0), ("b11.expand", 20 * 15, 80, 480, 2, 0, 0), ("b11.project", 20 * 15, 480, 112, 0, 0, 1),
     ("b12.expand", 20 * 15, 112, 672, 2, 0, 0), ("b12.project", 20 * 15, 672, 112, 0, 1, 1), ("b13.expand", 20 * 15, 112, 672, 2, 0, 0),
     ("b13.project", 20 * 15, 672, 160, 0, 0, 1), ("b14.expand", 20 * 15, 160, 960, 2, 0, 0), ("b14.project", 20 * 15, 960, 160, 0, 1, 1),
     ("b15.expand", 20 * 15, 160, 960, 2, 0, 0), ("b15.project", 20 * 15, 960, 160, 0, 1, 1), ("b16.conv", 20 * 15, 160, 960, 2, 0, 0)]
dev = "cuda"
g = torch.Generator(device=dev).manual_seed(0)
res = {"batch": B, "layers": {}}
total = 0.0
only = os.environ.get("GEMM_ONLY")
for (name, hw, K, N, act, has_res, has_se) in L:
    if only and name not in only.split(","):
        continue
    M = B * hw
    nbuf = 3
    xs = [torch.randn(M, K, device=dev, generator=g).bfloat16() for _ in range(nbuf)]
    w = (torch.randn(N, K, device=dev, generator=g) * K ** -0.5).bfloat16()
    sc = torch.rand(N, device=dev, generator=g) + 0.5
    sh = torch.randn(N, device=dev, generator=g) * 0.1
    r = torch.randn(M, N, device=dev, generator=g).bfloat16() if has_res else None
    gate = torch.rand(B, K, device=dev, generator=g) if has_se else None
    for i in range(nbuf):
        out = D.conv1x1(xs[i], w, sc, sh, act, r, gate, hw if has_se else 0)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for i in range(REPS):
        D.conv1x1(xs[i % nbuf], w, sc, sh, act, r, gate, hw if has_se else 0)
    e1.record(); torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 1e3 / REPS
    algo = (M * K + M * N * (2 if has_res else 1) + N * K) * 2
    res["layers"][name] = {"us": us, "GB/s": algo / us / 1e3, "TFLOP/s": 2.0 * M * N * K / us / 1e6, "abs_sum": float(out.float().abs().sum().item())}
    total += us
    del xs, out, r
    torch.cuda.empty_cache()
res["total_us"] = total
print(f"total {total:.1f} us | " + " ".join(f"{k}={v['us']:.0f}" for k, v in res["layers"].items()))
if len(sys.argv) > 1:
    json.dump(res, open(sys.argv[1], "w"), indent=1)
