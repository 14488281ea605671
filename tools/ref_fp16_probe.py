"""Diagnostic: where does the unmodified reference model first produce non-finite values under fp16 autocast (its AMP dtype)?"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from oracle import lraspp_oracle as O, ref_loader as R

mods = R.load_reference(("config", "model", "utils"))
torch.manual_seed(0)
ref = mods["model"].create_model(2, pretrained=False).cuda().train()
x, m = O.synthetic_cards(16, seed=1000)
xc, mc = x.cuda(), m.cuda()
bad = []
def hook(name):
    def f(mod, inp, out):
        t = out if torch.is_tensor(out) else None
        if t is not None and not torch.isfinite(t).all() and len(bad) < 5:
            bad.append((name, type(mod).__name__, str(t.dtype), float(t.float().abs().nan_to_num(posinf=1e9).max())))
    return f
for n, mod in ref.named_modules():
    if len(list(mod.children())) == 0:
        mod.register_forward_hook(hook(n))
crit = mods["utils"].CombinedLoss()
for dt in (torch.float16, torch.bfloat16):
    bad.clear()
    with torch.autocast("cuda", dtype=dt):
        out = ref(xc)
        loss = crit(out, mc)
    print(dt, "loss", float(loss), "out finite", bool(torch.isfinite(out).all()), "absmax", float(out.float().abs().max()), "first bad:", bad[:3])
out = ref(xc); print("fp32 loss", float(crit(out, mc)), float(out.abs().max()))
stats = {}
def hook2(name):
    def f(mod, inp, out):
        if torch.is_tensor(out): stats[name] = float(out.float().abs().max())
    return f
for n, mod in ref.named_modules():
    if len(list(mod.children())) == 0:
        mod.register_forward_hook(hook2(n))
with torch.no_grad(): ref(xc)
top = sorted(stats.items(), key=lambda kv: -kv[1])[:8]
print("largest activations fp32:", top)
