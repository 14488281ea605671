"""The reference's OWN drivers running on this package: the unmodified ``train_epoch`` / ``validate_epoch``
(train/train.py:67-153) and ``ModelEvaluator.evaluate_dataset`` (train/evaluate.py:41-137) are imported from
``/root/reference`` (build container) or the byte-identical staged copy ``oracle/_ref`` (GPU box; oracle/make_ref.py), with
their bare-name siblings ``model`` and ``utils`` replaced by ``mtg_card_image_segmentation_b200.model`` / ``.utils`` --
INTEGRATION.md §1, executed.  Everything else of the loop is the reference's: ``torch.autocast('cuda')`` (fp16), ``GradScaler``,
``create_optimizer`` (torch.optim.AdamW), ``Config``, sklearn's confusion matrix."""
import hashlib
import json
import os
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.dirname(__file__))

from oracle import lraspp_oracle as O  # noqa: E402
from oracle import ref_loader as R  # noqa: E402

needs_ref = pytest.mark.skipif(R.ref_dir(staged_only=True) is None, reason="reference not staged in oracle/_ref (python oracle/make_ref.py)")


class _Loader(list):
    """A list of {'image','mask','filename'} batches with the `.dataset` attribute evaluate.py:52 prints."""

    def __init__(self, batches):
        super().__init__(batches)
        self.dataset = [None] * sum(b["image"].shape[0] for b in batches)


def _batches(n_batches, batch, seed, h=128, w=96):
    out = []
    for i in range(n_batches):
        x, m = O.synthetic_cards(batch, seed=seed + i, height=h, width=w)
        out.append({"image": x, "mask": m, "filename": [f"card_{seed + i}_{j}.png" for j in range(batch)]})
    return _Loader(out)


@needs_ref
def test_staged_reference_is_unmodified():
    """oracle/_ref holds byte-identical copies (sha256 manifest written at staging time; compared with the source when present)."""
    man = os.path.join(ROOT, "oracle", "_ref", "MANIFEST.json")
    if not os.path.exists(man):
        pytest.skip("oracle/_ref not staged (the reference is imported from /root/reference)")
    shas = json.load(open(man))["sha256"]
    for rel, sha in shas.items():
        with open(os.path.join(ROOT, "oracle", "_ref", rel), "rb") as f:
            assert hashlib.sha256(f.read()).hexdigest() == sha, rel
        src = os.path.join("/root/reference", rel)
        if os.path.exists(src):
            with open(src, "rb") as f:
                assert hashlib.sha256(f.read()).hexdigest() == sha, rel


@needs_ref
@pytest.mark.gpu
def test_reference_train_validate_evaluate_on_the_drop_in():
    import mtg_card_image_segmentation_b200 as M
    from mtg_card_image_segmentation_b200 import evaluate as our_eval, model as our_model, utils as our_utils
    mods = R.load_reference(("config", "train", "evaluate"), shim={"model": our_model, "utils": our_utils}, staged_only=True)
    try:
        ref_train, ref_eval, Config = mods["train"], mods["evaluate"], mods["config"].Config
        assert ref_train.create_model is our_model.create_model and ref_train.CombinedLoss is our_utils.CombinedLoss
        assert Config.USE_AMP is True and str(Config.DEVICE) == "cuda"
        dev = torch.device("cuda")
        torch.manual_seed(0)
        model = ref_train.create_model(num_classes=Config.NUM_CLASSES, pretrained=False).to(dev)
        assert isinstance(model, M.CardSegmentationModel)
        criterion = ref_train.CombinedLoss(dice_weight=Config.DICE_WEIGHT, ce_weight=Config.BCE_WEIGHT)
        optimizer = ref_train.create_optimizer(model, Config)       # train/train.py:155-171: torch.optim.AdamW, all params
        assert type(optimizer) is torch.optim.AdamW
        scaler = ref_train.GradScaler("cuda")                       # train/train.py:272
        train_loader, val_loader = _batches(6, 8, seed=100), _batches(2, 8, seed=900)
        lib = M._native.load()
        l0 = lib.mtgseg_launch_count()
        first = ref_train.train_epoch(model, train_loader, criterion, optimizer, scaler, dev, 1)
        for epoch in range(2, 5):
            last = ref_train.train_epoch(model, train_loader, criterion, optimizer, scaler, dev, epoch)
        assert lib.mtgseg_launch_count() - l0 > 4 * 6 * 100, "the CUDA training step did not run"
        keys = {"loss", "iou_background", "iou_card", "mean_iou", "dice_background", "dice_card", "mean_dice", "pixel_accuracy"}
        assert set(first) == keys and set(last) == keys
        print(f"reference train_epoch on the drop-in: loss {first['loss']:.4f} -> {last['loss']:.4f}, scale {scaler.get_scale():.0f}")
        assert last["loss"] < first["loss"] and all(map(lambda v: v == v, last.values()))
        assert scaler.get_scale() >= 1024.0  # fp16 autocast + GradScaler: the scale survived (no overflow cascade)
        val = ref_train.validate_epoch(model, val_loader, criterion, dev, 1)
        assert set(val) == keys and 0.0 <= val["pixel_accuracy"] <= 1.0
        # evaluate.py:41-100 -- fp32, no autocast (-> the fp32-exact path), sklearn confusion matrix over Python lists
        res = ref_eval.ModelEvaluator(model, dev, num_classes=2).evaluate_dataset(val_loader, criterion)
        cm = res["confusion_matrix"]
        sd = {k: v.detach().cpu() for k, v in model.state_dict().items()}
        want = torch.zeros(4, dtype=torch.int64)
        with torch.no_grad():
            for b in val_loader:
                want += O.confusion_counts(O.forward(sd, b["image"]), b["mask"])
        print("confusion matrix through the reference's evaluator:", cm.tolist(), "oracle:", want.reshape(2, 2).tolist())
        assert int(cm.sum()) == 2 * 8 * 128 * 96
        # the fp32 CPU oracle on the same trained weights: near-tie pixels may flip within the 1e-4 logit tolerance
        assert (torch.as_tensor(cm).reshape(-1) - want).abs().sum().item() <= 2e-3 * int(cm.sum())
        assert set(res["per_class_metrics"]) == {"background", "card"} and len(res["filenames"]) == 16
        # and this package's own evaluator (fused int64 counts) agrees with the reference's sklearn matrix bit for bit
        ours = our_eval.ModelEvaluator(model, dev).evaluate_dataset(val_loader, criterion)
        assert (ours["confusion_matrix"] == cm).all()
        for k in keys:
            assert abs(ours["basic_metrics"][k] - res["basic_metrics"][k]) <= 1e-6
    finally:
        R.unload()


@needs_ref
@pytest.mark.gpu
def test_parity_on_weights_trained_by_the_reference():
    """SURVEY.md §8c fixture C, produced by the REFERENCE: its own unmodified ``create_model`` + ``train_epoch`` (torchvision /
    cuDNN, torch AdamW; fp32, see below) train the network on synthetic cards on this GPU; the resulting state_dict
    goes into this package with ``load_state_dict(strict=True)``.  North-star tolerances on those weights at config.py
    resolution, against the reference model's own fp32 forward: tensor-core path logits <= 2e-2 of the logit range and rel-L2,
    fp32-exact path <= 1e-4, thresholded masks >= 99.9 % identical over ALL pixels (both paths).
    (The 16.8 MB trained state_dict is not committed as a golden file: the staged reference regenerates it in ~20 s.)"""
    import mtg_card_image_segmentation_b200 as M
    import devops as D
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    mods = R.load_reference(("config", "model", "utils", "train"), staged_only=True)
    try:
        ref_train, Config = mods["train"], mods["config"].Config
        dev = torch.device("cuda")
        # The fixture is trained in fp32 (Config.USE_AMP = False: a configuration switch of the reference, train/config.py:32, not an
        # edit).  Measured (gpurun_out/r2_bisect.log, r2_t3.log): the reference's own fp16-autocast loop -- cuDNN / ATen only, no kernel
        # of this repository involved -- diverges to NaN on these synthetic cards in some processes (from the first batch when the
        # allocator's recycled blocks hold bf16 data of earlier tests, after a few epochs otherwise) and trains fine in others;
        # a golden fixture must not depend on that.  The drop-in test above exercises the fp16 + GradScaler loop on THIS package.
        Config.USE_AMP = False
        torch.cuda.synchronize()
        torch.cuda.empty_cache()
        z = torch.zeros(1 << 30, dtype=torch.float32, device=dev)  # recycled pages come back zeroed, not with stale bf16 patterns
        del z
        torch.manual_seed(0)
        ref = mods["model"].create_model(num_classes=2, pretrained=False).to(dev)
        crit = mods["utils"].CombinedLoss(dice_weight=Config.DICE_WEIGHT, ce_weight=Config.BCE_WEIGHT)
        opt = ref_train.create_optimizer(ref, Config)
        scaler = ref_train.GradScaler("cuda")
        loader = _batches(24, 16, seed=1000, h=Config.INPUT_HEIGHT, w=Config.INPUT_WIDTH)
        first = ref_train.train_epoch(ref, loader, crit, opt, scaler, dev, 1)
        for epoch in range(2, 9):
            last = ref_train.train_epoch(ref, loader, crit, opt, scaler, dev, epoch)
        print(f"reference train_epoch x8 (192 steps, B=16): loss {first['loss']:.4f} -> {last['loss']:.4f}, mean IoU {last['mean_iou']:.3f}")
        assert last["loss"] < 0.5 * first["loss"]
        # 192 steps at BatchNorm momentum 0.01 leave the running statistics far from converged: re-estimate them on a calibration
        # batch with the reference model itself (momentum=None = cumulative average), like a longer run would have
        ref.train()
        for mod in ref.modules():
            if isinstance(mod, torch.nn.BatchNorm2d):
                mod.reset_running_stats()
                mod.momentum = None
        xcal, _ = O.synthetic_cards(16, seed=77)
        with torch.no_grad():
            ref(xcal.to(dev))
        ref.eval()
        sd = {k: v.detach().clone() for k, v in ref.state_dict().items()}
        x, m = O.synthetic_cards(8, seed=4242)
        xc = x.to(dev)
        with torch.no_grad():
            zref = ref(xc).float().cpu()  # the reference's own fp32 forward (cuDNN, TF32 off)
            zcpu = O.forward({k: v.cpu() for k, v in sd.items()}, x)
        D.report("reference fp32 on cuDNN vs CPU oracle", zref, zcpu)
        ours = M.create_model(2, pretrained=False)
        ours.load_state_dict(sd, strict=True)
        ours = ours.to(dev).eval()
        with torch.no_grad():
            z32 = ours(xc).cpu()
            zbf = ours.predict(xc, want_logits=True, precision="bf16")["logits"].cpu()
    finally:
        R.unload()
    iou = O.metrics_from_counts(O.confusion_counts(zcpu, m))["iou"]
    print(f"IoU of the reference-trained network vs ground truth: {iou}")
    assert min(iou) > 0.5, "degenerate fixture: the reference-trained network must actually segment the cards"
    for name, z, tol in (("fp32-exact path", z32, 1e-4), ("tensor-core bf16 path", zbf, 2e-2)):
        for rname, r in (("reference fp32 (cuDNN)", zref), ("CPU oracle", zcpu)):
            emax, el2 = D.report(f"reference-trained weights: {name} vs {rname}", z, r)
            agree = ((z[:, 1] > z[:, 0]) == (r[:, 1] > r[:, 0])).float().mean().item()
            print(f"    mask agreement over all pixels: {agree:.6f}")
            # cuDNN's own fp32 differs from the CPU's by ~1e-6; the 1e-4 bar is asserted against both
            assert emax <= tol and el2 <= tol, (name, rname, emax, el2)
            assert agree >= 0.999, (name, rname, agree)
