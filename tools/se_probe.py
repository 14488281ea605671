#!/usr/bin/env python
"""Runs the squeeze-excite MLP of the nine SE blocks of the B=256 forward in isolation (timing, or as an ncu target)."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import devops as D
B = 256
dev = "cuda"
g = torch.Generator(device=dev).manual_seed(0)
only = os.environ.get("SE_ONLY")
for (name, C, SQ, chunks, hw) in [("b4", 72, 24, 10, 1200), ("b5", 120, 32, 5, 1200), ("b11", 480, 120, 1, 300), ("b12", 672, 168, 1, 300),
                                  ("b13", 672, 168, 2, 300), ("b14", 960, 240, 2, 300)]:
    if only and name != only:
        continue
    sums = torch.randn(B, chunks, C, device=dev, generator=g)
    w1 = (torch.randn(SQ, C, device=dev, generator=g) * C ** -0.5).bfloat16(); b1 = torch.randn(SQ, device=dev, generator=g) * 0.1
    w2 = (torch.randn(C, SQ, device=dev, generator=g) * SQ ** -0.5).bfloat16(); b2 = torch.randn(C, device=dev, generator=g) * 0.1
    for _ in range(3):
        D.se_mlp(sums, hw, w1, b1, 1, w2, b2, 3)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    for _ in range(10):
        D.se_mlp(sums, hw, w1, b1, 1, w2, b2, 3)
    e1.record(); torch.cuda.synchronize()
    print(f"{name}: C={C} SQ={SQ} chunks={chunks}: {e0.elapsed_time(e1) * 100:.1f} us")
