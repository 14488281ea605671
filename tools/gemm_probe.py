#!/usr/bin/env python
"""Runs three representative conv GEMM launches (for ncu captures): b16.conv 160->960, b2.expand 16->64, head 3x3."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import devops as D
B = 256
dev = "cuda"
def run(M, N, K, act, iters=3):
    x = torch.randn(M, K, device=dev).bfloat16(); w = torch.randn(N, K, device=dev).bfloat16()
    sc = torch.ones(N, device=dev); sh = torch.zeros(N, device=dev)
    for _ in range(iters):
        D.conv1x1(x, w, sc, sh, act)
    torch.cuda.synchronize()
run(B * 300, 960, 160, 2)
run(B * 19200, 64, 16, 1)
run(B * 4800, 24, 72, 0)
x = torch.randn(B, 20, 15, 960, device=dev).bfloat16(); w = (torch.randn(128, 9, 960, device=dev) * 0.01).bfloat16()
sc = torch.ones(128, device=dev); sh = torch.zeros(128, device=dev)
for _ in range(3):
    D.conv3x3(x, w, sc, sh, 1)
torch.cuda.synchronize()
print("ok")
