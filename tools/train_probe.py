#!/usr/bin/env python
"""Runs a few B=32 training steps through the public surface (for ncu launch lists / nsight captures)."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import mtg_card_image_segmentation_b200 as M
from mtg_card_image_segmentation_b200.optim import FusedAdamW
B = int(os.environ.get("TRAIN_B", "32"))
steps = int(os.environ.get("TRAIN_STEPS", "3"))
g = torch.Generator().manual_seed(1)  # noise images with a rectangular "card" mask (the oracle is test-only: not imported here)
x = torch.randn(B, 3, 320, 240, generator=g).cuda()
m = torch.zeros(B, 320, 240, dtype=torch.int64)
m[:, 60:260, 50:190] = 1
m = m.cuda()
torch.manual_seed(0)
model = M.create_model(2, False).cuda().train()
opt = FusedAdamW(model.parameters(), lr=1e-3, weight_decay=1e-4)
crit = M.CombinedLoss()
import time
for i in range(steps):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    opt.zero_grad(set_to_none=True)
    loss = crit(model(x), m)
    loss.backward()
    opt.step()
    torch.cuda.synchronize()
    print(f"step {i}: loss {loss.item():.4f} wall {1e3*(time.perf_counter()-t0):.2f} ms")
