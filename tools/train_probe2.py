#!/usr/bin/env python
"""Splits a B=32 training step into host-enqueue time vs device time per phase."""
import os, sys, time
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import mtg_card_image_segmentation_b200 as M
from mtg_card_image_segmentation_b200.optim import FusedAdamW
B = int(os.environ.get("TRAIN_B", "32"))
g = torch.Generator().manual_seed(1)  # noise images with a rectangular "card" mask (the oracle is test-only: not imported here)
x = torch.randn(B, 3, 320, 240, generator=g).cuda()
m = torch.zeros(B, 320, 240, dtype=torch.int64)
m[:, 60:260, 50:190] = 1
m = m.cuda()
model = M.create_model(2, False).cuda().train()
opt = FusedAdamW(model.parameters(), lr=1e-3, weight_decay=1e-4)
crit = M.CombinedLoss()
names = ["zero_grad", "forward", "loss", "backward", "opt.step"]
for it in range(6):
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(6)]
    host = []
    torch.cuda.synchronize()
    t = time.perf_counter(); ev[0].record()
    opt.zero_grad(set_to_none=True); host.append(time.perf_counter() - t); ev[1].record(); t = time.perf_counter()
    out = model(x); host.append(time.perf_counter() - t); ev[2].record(); t = time.perf_counter()
    loss = crit(out, m); host.append(time.perf_counter() - t); ev[3].record(); t = time.perf_counter()
    loss.backward(); host.append(time.perf_counter() - t); ev[4].record(); t = time.perf_counter()
    opt.step(); host.append(time.perf_counter() - t); ev[5].record()
    torch.cuda.synchronize()
    if it >= 3:
        print(" | ".join(f"{n}: host {1e3*h:.2f} dev {ev[i].elapsed_time(ev[i+1]):.2f}" for i, (n, h) in enumerate(zip(names, host))),
              f"| total dev {ev[0].elapsed_time(ev[5]):.2f} ms")
