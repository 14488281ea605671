#!/usr/bin/env python
"""Times and checks the multi-tap tcgen05 conv (head 3x3: 960 -> 128 at 20x15) through the C ABI."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import torch.nn.functional as F
import devops as D

torch.backends.cudnn.allow_tf32 = False
for (B, H, W, K, N) in [(2, 4, 3, 64, 32), (3, 20, 15, 960, 128), (256, 20, 15, 960, 128), (32, 20, 15, 128, 960)]:
    g = torch.Generator().manual_seed(B + K)
    a = torch.randn(B, H, W, K, generator=g).to(torch.bfloat16).cuda()
    w = (torch.randn(N, 9, K, generator=g) / (9 * K) ** 0.5).to(torch.bfloat16).cuda()
    scale = (torch.rand(N, generator=g) + 0.5).cuda(); shift = torch.randn(N, generator=g).cuda()
    out = D.conv3x3(a, w, scale, shift, act=1)
    ref = F.conv2d(a.float().permute(0, 3, 1, 2), w.float().view(N, 3, 3, K).permute(0, 3, 1, 2), padding=1)
    ref = F.relu(ref * scale.view(1, -1, 1, 1) + shift.view(1, -1, 1, 1)).permute(0, 2, 3, 1)
    err = float((out.float() - ref).abs().max() / ref.abs().max())
    for _ in range(5):
        D.conv3x3(a, w, scale, shift, act=1)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    for _ in range(20):
        D.conv3x3(a, w, scale, shift, act=1)
    e1.record(); torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / 20 * 1e3
    print(f"B={B} {H}x{W} {K}->{N}: max err {err:.2e}  {us:8.1f} us  {2.0 * B * H * W * N * K * 9 / us / 1e6:7.1f} TFLOP/s")
