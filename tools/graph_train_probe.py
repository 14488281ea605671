#!/usr/bin/env python
"""Debug aid for engine.GraphedTrainStep: builds the captured training step at several batch shapes and prints the stream-capture
status after every C-ABI call made during the capture (which call invalidated it, if any).

  python tools/graph_train_probe.py "4,64,48 32,320,240"
"""
import os, sys, traceback
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import mtg_card_image_segmentation_b200 as M
from mtg_card_image_segmentation_b200 import _native as N
from mtg_card_image_segmentation_b200.optim import FusedAdamW
from mtg_card_image_segmentation_b200.engine import GraphedTrainStep
from cuda.bindings import runtime as rt

lib = N.load()
log = []


def status():
    err, st = rt.cudaStreamIsCapturing(torch.cuda.current_stream().cuda_stream)
    return f"{err.name}/{st.name}"


def wrap(name):
    orig = getattr(lib, name)

    def f(*a):
        rc = orig(*a)
        s = status()
        if "None" not in s:  # only while capturing
            log.append(f"{name} rc={rc} -> {s}")
        return rc
    setattr(lib, name, f)


for n in ("mtgseg_pack_weights", "mtgseg_forward_train", "mtgseg_loss_fwd_bwd", "mtgseg_backward", "mtgseg_adamw_step_dev"):
    wrap(n)

if len(sys.argv) > 1 and sys.argv[1] == "--bench":  # run bench.py with the wrappers installed; print the capture log at exit
    import atexit, runpy
    atexit.register(lambda: print("   " + "\n   ".join(log), file=sys.stderr, flush=True))
    sys.argv = ["bench.py"] + sys.argv[2:]
    runpy.run_path(os.path.join(ROOT, "bench.py"), run_name="__main__")
    sys.exit(0)

shapes = [tuple(int(v) for v in s.split(",")) for s in (sys.argv[1] if len(sys.argv) > 1 else "4,64,48 32,320,240").split()]
for (B, H, W) in shapes:
    torch.manual_seed(0)
    model = M.create_model(2, pretrained=False).cuda().train()
    opt = FusedAdamW(model.parameters(), lr=1e-3, weight_decay=1e-4)
    crit = M.CombinedLoss()
    x = torch.randn(B, 3, H, W, device="cuda")
    y = (torch.rand(B, H, W, device="cuda") > 0.5).long()
    if os.environ.get("PROBE_EAGER_FIRST", "0") == "1":
        for _ in range(2):
            opt.zero_grad(set_to_none=True)
            crit(model(x), y).backward()
            opt.step()
        torch.cuda.synchronize()
    del log[:]
    try:
        g = GraphedTrainStep(model, crit, opt, x, y)
        l = [float(g.step(x, y)) for _ in range(3)]
        print(f"B={B} {H}x{W}: OK {g.launches_per_replay} launches, losses {l}", flush=True)
    except Exception as e:
        print(f"B={B} {H}x{W}: FAILED {type(e).__name__}: {str(e).splitlines()[0]}", flush=True)
        tb = traceback.format_exc().splitlines()
        print("\n".join(tb[-12:]), flush=True)
    print("   " + "\n   ".join(log), flush=True)
    del model, opt
    torch.cuda.synchronize()
    torch.cuda.empty_cache()
