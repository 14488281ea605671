"""TEST INFRASTRUCTURE ONLY — generates tests/golden/*.pt from the UNMODIFIED reference.

Run in the build container (needs /root/reference and torchvision 0.26.0):

    python oracle/make_golden.py

It imports the reference's own ``train/model.py`` / ``train/utils.py`` with the two
harness-side shims of SURVEY.md App. A (no ImageNet download; stub modules for the plotting /
augmentation wheels that are absent), loads deterministic weights (``oracle.make_weights``)
with ``load_state_dict(strict=True)``, and records what the reference computes:

  seg_small.pt   64x48 inputs, B=2: eval logits, train-mode logits + BN updates, loss, selected
                 parameter gradients, a couple of named activations
  seg_full.pt    320x240 (config.py resolution), B=1: eval logits (fp16-packed diff-safe copy in fp32),
                 low-res logits, loss, confusion counts, calibrated running stats
  metrics.pt     MetricsCalculator / calculate_* outputs on seeded random logits incl. ties,
                 per-class metrics of evaluate.py on an integer confusion matrix
  adamw.pt       torch.optim.AdamW two-step trajectory on a small tensor
"""
import functools
import os
import sys
import types

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import lraspp_oracle as O  # noqa: E402

REF = "/root/reference/train"


def import_reference():
    for n in ("matplotlib", "matplotlib.pyplot", "seaborn", "albumentations", "albumentations.pytorch"):
        sys.modules.setdefault(n, types.ModuleType(n))
    sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
    sys.modules["albumentations.pytorch"].ToTensorV2 = object
    sys.path.insert(0, REF)
    import model as ref_model
    ref_model.lraspp_mobilenet_v3_large = functools.partial(ref_model.lraspp_mobilenet_v3_large, weights_backbone=None)
    import utils as ref_utils
    import evaluate as ref_eval
    return ref_model, ref_utils, ref_eval


sample = O.sample


def main():
    torch.manual_seed(0)
    torch.set_num_threads(8)
    ref_model, ref_utils, ref_eval = import_reference()
    out_dir = os.path.join(ROOT, "tests", "golden")
    os.makedirs(out_dir, exist_ok=True)

    model = ref_model.create_model(num_classes=2, pretrained=False)
    crit = ref_utils.CombinedLoss(dice_weight=0.5, ce_weight=0.5)

    # ---------------- small resolution, B=2 -------------------------------------------------
    xs, ms = O.synthetic_cards(2, seed=7, height=64, width=48)
    sd = O.calibrate_running_stats(O.make_weights(11), xs)
    rstats = {k: v for k, v in sd.items() if "running_" in k}
    model.load_state_dict(sd, strict=True)
    assert list(model.state_dict().keys()) == [k for k, _, _ in O.state_dict_spec()]
    model.eval()
    acts = {}
    hooks = [
        model.model.backbone["0"].register_forward_hook(lambda m, i, o: acts.__setitem__("stem", o.detach().clone())),
        model.model.backbone["4"].register_forward_hook(lambda m, i, o: acts.__setitem__("b4.out", o.detach().clone())),
        model.model.backbone["16"].register_forward_hook(lambda m, i, o: acts.__setitem__("high", o.detach().clone())),
        model.model.classifier.register_forward_hook(lambda m, i, o: acts.__setitem__("lowres_logits", o.detach().clone())),
    ]
    with torch.no_grad():
        eval_logits = model(xs)
    eval_acts = dict(acts)
    for h in hooks:
        h.remove()
    model.train()
    model.zero_grad()
    train_logits = model(xs)
    loss = crit(train_logits, ms)
    loss.backward()
    after = model.state_dict()
    grads = {k: p.grad.detach().clone() for k, p in model.named_parameters()}
    grad_keys = ["model.backbone.0.0.weight", "model.backbone.0.1.weight", "model.backbone.4.block.2.fc1.weight",
                 "model.backbone.4.block.2.fc2.bias", "model.backbone.15.block.1.0.weight",
                 "model.backbone.15.block.3.0.weight", "model.backbone.16.1.bias", "model.classifier.cbr.0.weight",
                 "model.classifier.scale.1.weight", "model.classifier.low_classifier.weight",
                 "model.classifier.high_classifier.bias"]
    torch.save({
        "weights_seed": 11, "input_seed": 7, "height": 64, "width": 48, "batch": 2,
        "running_stats": rstats,
        "eval_logits": eval_logits, "eval_acts": eval_acts,
        "train_logits": train_logits.detach(), "train_loss": loss.detach(),
        "bn_after": {k: after[k].clone() for k in ("model.backbone.0.1.running_mean", "model.backbone.0.1.running_var",
                                                   "model.backbone.0.1.num_batches_tracked",
                                                   "model.classifier.cbr.1.running_mean", "model.classifier.cbr.1.running_var",
                                                   "model.backbone.15.block.1.1.running_var")},
        "grads": {k: (grads[k] if grads[k].numel() <= 8192 else sample(grads[k], 4096)) for k in grad_keys},
        "grad_norms": {k: v.norm() for k, v in grads.items()},
    }, os.path.join(out_dir, "seg_small.pt"))

    # ---------------- config.py resolution, B=1 ---------------------------------------------
    xf, mf = O.synthetic_cards(4, seed=1234)
    sdf = O.calibrate_running_stats(O.make_weights(3), xf)
    model.load_state_dict(sdf, strict=True)
    model.eval()
    acts.clear()
    h = model.model.classifier.register_forward_hook(lambda m, i, o: acts.__setitem__("lowres_logits", o.detach().clone()))
    with torch.no_grad():
        lf = model(xf[:1])
        lossf = crit(lf, mf[:1])
    h.remove()
    pred = torch.argmax(lf, 1)
    cm = torch.zeros(2, 2, dtype=torch.int64)
    for t in range(2):
        for p in range(2):
            cm[t, p] = ((mf[:1] == t) & (pred == p)).sum()
    torch.save({
        "weights_seed": 3, "input_seed": 1234, "calib_batch": 4, "height": 320, "width": 240, "batch": 1,
        "running_stats": {k: v for k, v in sdf.items() if "running_" in k},
        "lowres_logits": acts["lowres_logits"], "logits_sample": sample(lf, 4096),
        "logits_sum": lf.double().sum(), "logits_abs_sum": lf.double().abs().sum(),
        "mask_u8": pred.to(torch.uint8),
        "loss": lossf, "confusion": cm,
    }, os.path.join(out_dir, "seg_full.pt"))

    # ---------------- metrics ----------------------------------------------------------------
    g = torch.Generator().manual_seed(5)
    cases = []
    for (b, hh, ww) in [(2, 64, 48), (3, 17, 5), (1, 320, 240)]:
        z = torch.randn(b, 2, hh, ww, generator=g)
        z[:, 1, ::3, ::2] = z[:, 0, ::3, ::2]  # exact ties -> argmax picks class 0
        t = torch.randint(0, 2, (b, hh, ww), generator=g)
        mc = ref_utils.MetricsCalculator(num_classes=2, device="cpu")
        l = crit(z, t)
        mc.update(l, z, t)
        mc.update(l * 0.5, z.flip(0), t)
        cases.append({"logits": z if hh < 100 else None, "seed_shape": (b, hh, ww), "targets": t if hh < 100 else None,
                      "loss": l, "iou": ref_utils.calculate_iou(z, t), "dice": ref_utils.calculate_dice_coefficient(z, t),
                      "acc": ref_utils.calculate_pixel_accuracy(z, t), "epoch_metrics": mc.get_metrics(),
                      "counts": O.confusion_counts(z, t)})
    ev = ref_eval.ModelEvaluator(torch.nn.Identity(), "cpu", 2)
    import numpy as np
    cm_np = np.array([[123456, 789], [1011, 98765]], dtype=np.int64)
    pcm = ev._calculate_per_class_metrics(cm_np)
    torch.save({"cases": cases, "cm": torch.from_numpy(cm_np),
                "per_class": {k: {kk: float(vv) for kk, vv in v.items()} for k, v in pcm.items()}},
               os.path.join(out_dir, "metrics.pt"))

    # ---------------- AdamW -----------------------------------------------------------------
    g = torch.Generator().manual_seed(9)
    p = torch.nn.Parameter(torch.randn(257, generator=g))
    opt = torch.optim.AdamW([p], lr=1e-3, weight_decay=1e-4)
    traj = {"p0": p.detach().clone(), "g": [], "p": []}
    for _ in range(3):
        p.grad = torch.randn(257, generator=g)
        traj["g"].append(p.grad.clone())
        opt.step()
        traj["p"].append(p.detach().clone())
    torch.save(traj, os.path.join(out_dir, "adamw.pt"))
    for f in sorted(os.listdir(out_dir)):
        print(f, os.path.getsize(os.path.join(out_dir, f)))


if __name__ == "__main__":
    main()
