"""Diagnostic: the GradScaler-vs-plain step of tests/test_gpu_train.py with per-parameter reporting."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from oracle import lraspp_oracle as O
import mtg_card_image_segmentation_b200 as M
from mtg_card_image_segmentation_b200.optim import FusedAdamW

x, m = O.synthetic_cards(4, seed=5, height=64, width=48)
sd = O.make_weights(41)
xc, mc = x.cuda(), m.cuda()

def one_step(scale):
    model = M.create_model(2, False); model.load_state_dict(sd); model = model.cuda().train()
    opt = FusedAdamW(model.parameters(), lr=1e-3, weight_decay=1e-4)
    crit = M.CombinedLoss()
    loss = crit(model(xc), mc)
    if scale:
        scaler = torch.amp.GradScaler("cuda", init_scale=scale)
        scaler.scale(loss).backward()
        scaler.unscale_(opt)
        g = {n: p.grad.detach().clone() for n, p in model.named_parameters()}
        scaler.step(opt); scaler.update()
    else:
        loss.backward()
        g = {n: p.grad.detach().clone() for n, p in model.named_parameters()}
        opt.step()
    torch.cuda.synchronize()
    return g, {n: p.detach().clone() for n, p in model.named_parameters()}

for trial in range(2):
    g0, p0 = one_step(0)
    g1, p1 = one_step(1024.0)
    rows = []
    for n in p0:
        d = (p0[n] - p1[n]).abs()
        i = int(d.argmax())
        rows.append((float(d.max()), n, float(g0[n].flatten()[i]), float(g1[n].flatten()[i]), int((d > 1e-4).sum()), p0[n].numel()))
    rows.sort(reverse=True)
    print("trial", trial)
    for r_ in rows[:6]:
        print(f"   dparam {r_[0]:.3e} {r_[1]:48s} g_plain {r_[2]:.3e} g_scaled {r_[3]:.3e}  elems>1e-4: {r_[4]}/{r_[5]}")
