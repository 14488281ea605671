#!/usr/bin/env python
"""Runs a few B=32 training steps through the public surface (for ncu launch lists / nsight captures)."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import mtg_card_image_segmentation_b200 as M
from mtg_card_image_segmentation_b200.optim import FusedAdamW
from oracle.lraspp_oracle import synthetic_cards
B = int(os.environ.get("TRAIN_B", "32"))
steps = int(os.environ.get("TRAIN_STEPS", "3"))
x, m = synthetic_cards(min(B, 8), seed=1)
x = x.repeat((B + 7) // 8, 1, 1, 1)[:B].cuda(); m = m.repeat((B + 7) // 8, 1, 1)[:B].cuda()
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
