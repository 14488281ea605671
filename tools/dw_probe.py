#!/usr/bin/env python
"""Runs the b14 (5x5 dil 2, C=960, 20x15) and b3 (3x3, C=72, 80x60) depthwise launches (for ncu captures)."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import devops as D
B = 256
dev = "cuda"
for (H, W, C, k, s, d, gap) in [(20, 15, 960, 5, 1, 2, True), (80, 60, 72, 3, 1, 1, False)]:
    x = torch.randn(B, H, W, C, device=dev).bfloat16(); w = torch.randn(k * k, C, device=dev).bfloat16()
    sc = torch.ones(C, device=dev); sh = torch.zeros(C, device=dev)
    for _ in range(2):
        D.dwconv(x, w, sc, sh, 2, k, s, d, gap)
    torch.cuda.synchronize()
print("ok")
