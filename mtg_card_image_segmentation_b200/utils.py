"""``train/utils.py`` surface (losses, metrics, checkpoints) on top of the sm_100a kernels.

CUDA tensors go through ``libmtgseg_b200.so``: the loss and its gradient are ONE fused pass
(``mtgseg_loss_fwd_bwd``), every IoU/Dice/accuracy number derives from ONE integer reduction
(``mtgseg_metric_counts``).  CPU tensors are rejected — there is no fallback path.
"""
from __future__ import annotations

import os

import torch
import torch.nn as nn

from . import _native as N

_DT = {torch.float32: N.LOGITS_F32, torch.bfloat16: N.LOGITS_BF16, torch.float16: N.LOGITS_F16}


def _require_cuda(t, what):
    if not t.is_cuda:
        raise RuntimeError(f"{what}: CUDA tensors required (the B200 build has no CPU fallback)")


def _check_pair(predictions, targets, what):
    _require_cuda(predictions, what)
    if predictions.dim() != 4 or targets.dim() != 3 or predictions.shape[0] != targets.shape[0] \
            or predictions.shape[2:] != targets.shape[1:]:
        raise RuntimeError(f"{what}: expected predictions (B,C,H,W) and targets (B,H,W), got "
                           f"{tuple(predictions.shape)} / {tuple(targets.shape)}")
    if targets.dtype != torch.int64:
        raise RuntimeError(f"{what}: targets must be int64 (train/dataset.py:84-88)")
    if predictions.dtype not in _DT:
        raise RuntimeError(f"{what}: unsupported logits dtype {predictions.dtype}")


def fused_loss(predictions, targets, dice_weight, ce_weight, smooth, need_grad):
    """One kernel pass: loss3 = (total, dice, ce) and, when asked, d(total)/d(logits).  Also what engine.GraphedTrainStep calls
    directly (no autograd inside a captured step)."""
    lib = N.load()
    p = predictions.contiguous()
    t = targets.contiguous()
    B, C, H, W = p.shape
    # fp16 logits (the reference's autocast dtype) get an fp32 gradient: unscaled it is ~1e-7, an fp16 subnormal; it is narrowed
    # only after GradScaler's factor has been multiplied in (_FusedLoss.backward), like autocast's fp32 softmax / CE backward
    dlogits = None
    if need_grad:
        dlogits = torch.empty_like(p, dtype=torch.float32) if p.dtype == torch.float16 else torch.empty_like(p)
    scratch = torch.empty(lib.mtgseg_loss_scratch_bytes() // 4, dtype=torch.float32, device=p.device)
    loss3 = torch.empty(3, dtype=torch.float32, device=p.device)
    with torch.cuda.device(p.device):
        N.check(lib.mtgseg_loss_fwd_bwd(p.data_ptr(), _DT[p.dtype], t.data_ptr(), N.ptr(dlogits),
                                        _DT[dlogits.dtype] if need_grad else N.LOGITS_NONE, scratch.data_ptr(),
                                        loss3.data_ptr(), B, H * W, C, dice_weight, ce_weight, smooth, N.stream_ptr()),
                "mtgseg_loss_fwd_bwd")
    return loss3, dlogits


def loss_weights(criterion):
    """(dice_weight, ce_weight, smooth) of a DiceLoss / CombinedLoss module."""
    if isinstance(criterion, CombinedLoss):
        return float(criterion.dice_weight), float(criterion.ce_weight), float(criterion.dice_loss.smooth)
    if isinstance(criterion, DiceLoss):
        return 1.0, 0.0, float(criterion.smooth)
    raise RuntimeError("expected this package's CombinedLoss or DiceLoss (train/utils.py:15-92)")


class _FusedLoss(torch.autograd.Function):
    """loss3 = (total, dice, ce); the gradient w.r.t. the logits is produced by the same kernel pass."""

    @staticmethod
    def forward(ctx, predictions, targets, dice_weight, ce_weight, smooth):
        loss3, dlogits = fused_loss(predictions, targets, dice_weight, ce_weight, smooth, predictions.requires_grad)
        ctx.dlogits = dlogits
        ctx.out_dtype = predictions.dtype
        ctx.mark_non_differentiable(loss3)
        return loss3[0], loss3

    @staticmethod
    def backward(ctx, grad_total, _grad_parts):
        if ctx.dlogits is None:
            return None, None, None, None, None
        # scale in the gradient's own precision (fp32 for fp16 logits), narrow afterwards (autograd casts to the logits' dtype too)
        return (ctx.dlogits * grad_total.to(ctx.dlogits.dtype)).to(ctx.out_dtype), None, None, None, None


class DiceLoss(nn.Module):
    """Global soft-Dice over batch, classes and pixels (train/utils.py:15-56)."""

    def __init__(self, smooth=1e-6):
        super().__init__()
        self.smooth = smooth

    def forward(self, predictions, targets):
        _check_pair(predictions, targets, "DiceLoss")
        return _FusedLoss.apply(predictions, targets, 1.0, 0.0, float(self.smooth))[0]


class CombinedLoss(nn.Module):
    """dice_weight * DiceLoss + ce_weight * CrossEntropyLoss (train/utils.py:58-92)."""

    def __init__(self, dice_weight=0.5, ce_weight=0.5, class_weights=None):
        super().__init__()
        if class_weights is not None:
            raise RuntimeError("class_weights is not supported by the fused loss kernel (the reference's drivers never "
                               "pass it: train/train.py:260, train/evaluate.py:385)")
        self.dice_weight = dice_weight
        self.ce_weight = ce_weight
        self.dice_loss = DiceLoss()
        self.ce_loss = nn.CrossEntropyLoss(weight=None)  # kept for attribute parity; the kernel computes CE itself

    def forward(self, predictions, targets):
        _check_pair(predictions, targets, "CombinedLoss")
        return _FusedLoss.apply(predictions, targets, float(self.dice_weight), float(self.ce_weight),
                                float(self.dice_loss.smooth))[0]


# ------------------------------------------------------------------------------------------------
# metrics
# ------------------------------------------------------------------------------------------------

def confusion_counts(predictions, targets, out=None):
    """int64[4] device tensor [n00, n01, n10, n11] (index = target*2 + argmax prediction, ties -> class 0).
    Accumulates into ``out`` when given.  The integer core of train/utils.py:94-164 and evaluate.py:88."""
    _check_pair(predictions, targets, "confusion_counts")
    if predictions.shape[1] != 2:
        raise RuntimeError("confusion_counts supports num_classes == 2 (train/config.py:20)")
    lib = N.load()
    p, t = predictions.contiguous(), targets.contiguous()
    if out is None:
        out = torch.zeros(4, dtype=torch.int64, device=p.device)
    B, _, H, W = p.shape
    with torch.cuda.device(p.device):
        N.check(lib.mtgseg_metric_counts(p.data_ptr(), _DT[p.dtype], t.data_ptr(), out.data_ptr(), B, H * W, N.stream_ptr()),
                "mtgseg_metric_counts")
    return out


def _ratios(counts, smooth):
    """Per-class intersection / prediction / target sums as float32 device scalars (what the reference's
    ``.float()`` mask sums give, exactly, below 2**24 pixels)."""
    c = counts.to(torch.float32)
    inter = torch.stack([c[0], c[3]])
    pred = torch.stack([c[0] + c[2], c[1] + c[3]])
    tgt = torch.stack([c[0] + c[1], c[2] + c[3]])
    return inter, pred, tgt


def calculate_iou(predictions, targets, num_classes=2, smooth=1e-6, _counts=None):
    """Per-class IoU (train/utils.py:94-121)."""
    c = _counts if _counts is not None else confusion_counts(predictions, targets)
    inter, pred, tgt = _ratios(c, smooth)
    return (inter + smooth) / (pred + tgt - inter + smooth)


def calculate_dice_coefficient(predictions, targets, num_classes=2, smooth=1e-6, _counts=None):
    """Per-class Dice (train/utils.py:123-149)."""
    c = _counts if _counts is not None else confusion_counts(predictions, targets)
    inter, pred, tgt = _ratios(c, smooth)
    return (2.0 * inter + smooth) / (pred + tgt + smooth)


def calculate_pixel_accuracy(predictions, targets, _counts=None):
    """Pixel accuracy (train/utils.py:151-164)."""
    c = _counts if _counts is not None else confusion_counts(predictions, targets)
    c = c.to(torch.float32)
    return (c[0] + c[3]) / c.sum()


class MetricsCalculator:
    """Accumulates per-batch ratios exactly like train/utils.py:166-225 (epoch value = mean of per-batch
    values) but from one counts kernel per batch and with NO host sync in ``update``; the global integer
    confusion matrix (train/evaluate.py:88) accumulates alongside in ``total_counts``."""

    def __init__(self, num_classes=2, device="cpu"):
        self.num_classes = num_classes
        self.device = torch.device(device) if isinstance(device, str) else device
        self.reset()

    def reset(self):
        self.total_loss = 0.0
        self._loss_acc = None
        self.total_iou = torch.zeros(self.num_classes, device=self.device)
        self.total_dice = torch.zeros(self.num_classes, device=self.device)
        self.total_accuracy = 0.0
        self._acc_acc = None
        self.total_counts = None
        self.count = 0

    def update(self, loss, predictions, targets):
        counts = confusion_counts(predictions, targets)
        self.total_counts = counts.clone() if self.total_counts is None else self.total_counts + counts
        self.total_iou = self.total_iou + calculate_iou(None, None, self.num_classes, _counts=counts).to(self.device)
        self.total_dice = self.total_dice + calculate_dice_coefficient(None, None, self.num_classes, _counts=counts).to(self.device)
        acc = calculate_pixel_accuracy(None, None, _counts=counts)
        self._acc_acc = acc if self._acc_acc is None else self._acc_acc + acc
        l = loss.detach().float() if torch.is_tensor(loss) else torch.tensor(float(loss))
        self._loss_acc = l if self._loss_acc is None else self._loss_acc + l.to(self._loss_acc.device)
        self.count += 1

    def get_metrics(self):
        if self.count == 0:
            return {}
        self.total_loss = float(self._loss_acc.item())
        self.total_accuracy = float(self._acc_acc.item())
        n = self.count
        return {
            "loss": self.total_loss / n,
            "iou_background": self.total_iou[0].item() / n,
            "iou_card": self.total_iou[1].item() / n,
            "mean_iou": self.total_iou.mean().item() / n,
            "dice_background": self.total_dice[0].item() / n,
            "dice_card": self.total_dice[1].item() / n,
            "mean_dice": self.total_dice.mean().item() / n,
            "pixel_accuracy": self.total_accuracy / n,
        }

    def confusion_matrix(self):
        """Global int64 2x2 matrix cm[target, prediction] over every pixel seen (train/evaluate.py:88)."""
        if self.total_counts is None:
            return torch.zeros(2, 2, dtype=torch.int64)
        return self.total_counts.reshape(2, 2).cpu()


def per_class_metrics(cm):
    """precision / recall / f1 / iou / support per class from the integer confusion matrix, in Python
    float arithmetic like train/evaluate.py:102-137."""
    cm = [[int(cm[0][0]), int(cm[0][1])], [int(cm[1][0]), int(cm[1][1])]]
    out = {}
    for i, name in enumerate(("background", "card")):
        tp = cm[i][i]
        fp = cm[0][i] + cm[1][i] - tp
        fn = cm[i][0] + cm[i][1] - tp
        precision = tp / (tp + fp) if (tp + fp) > 0 else 0
        recall = tp / (tp + fn) if (tp + fn) > 0 else 0
        f1 = 2 * precision * recall / (precision + recall) if (precision + recall) > 0 else 0
        iou = tp / (tp + fp + fn) if (tp + fp + fn) > 0 else 0
        out[name] = {"precision": precision, "recall": recall, "f1": f1, "iou": iou, "support": cm[i][0] + cm[i][1]}
    return out


# ------------------------------------------------------------------------------------------------
# checkpoints (layout contract of train/utils.py:227-280)
# ------------------------------------------------------------------------------------------------

def save_checkpoint(model, optimizer, scheduler, epoch, best_metric, checkpoint_dir, filename):
    os.makedirs(checkpoint_dir, exist_ok=True)
    path = os.path.join(checkpoint_dir, filename)
    torch.save({
        "epoch": epoch,
        "model_state_dict": model.state_dict(),
        "optimizer_state_dict": optimizer.state_dict() if optimizer else None,
        "scheduler_state_dict": scheduler.state_dict() if scheduler else None,
        "best_metric": best_metric,
    }, path)
    print(f"Checkpoint saved: {path}")


def load_checkpoint(model, optimizer, scheduler, checkpoint_path):
    ckpt = torch.load(checkpoint_path, map_location="cpu")
    model.load_state_dict(ckpt["model_state_dict"])
    optimizer.load_state_dict(ckpt["optimizer_state_dict"])
    if scheduler and ckpt.get("scheduler_state_dict"):
        scheduler.load_state_dict(ckpt["scheduler_state_dict"])
    epoch = ckpt.get("epoch", 0)
    best_metric = ckpt.get("best_metric", 0.0)
    print(f"Checkpoint loaded from: {checkpoint_path}")
    print(f"Resumed from epoch: {epoch}, Best metric: {best_metric:.4f}")
    return epoch, best_metric


# ------------------------------------------------------------------------------------------------
# the remaining names train/train.py:18-21 and train/evaluate.py:17-20 import from `utils`, so that this module can stand in
# for it as a whole (tests/test_reference_drivers.py runs the reference's drivers that way)
# ------------------------------------------------------------------------------------------------

_PRINTED = (("Loss", "loss"), ("Mean IoU", "mean_iou"), ("IoU Background", "iou_background"), ("IoU Card", "iou_card"),
            ("Mean Dice", "mean_dice"), ("Dice Background", "dice_background"), ("Dice Card", "dice_card"),
            ("Pixel Accuracy", "pixel_accuracy"))


def print_metrics(metrics, prefix=""):
    """Same eight lines, labels and 4-decimal format as train/utils.py:399-415."""
    print(f"{prefix}Metrics:")
    for label, key in _PRINTED:
        print(f"  {label}: {metrics.get(key, 0):.4f}")


def _visualisation_only(name):
    def fn(*args, **kwargs):
        raise RuntimeError(f"{name} is matplotlib visualisation (train/utils.py:282-397), outside the B200 hot path: "
                           "call the reference's own helper (it works on this package's model and metrics unchanged).")
    fn.__name__ = name
    return fn


plot_training_history = _visualisation_only("plot_training_history")
visualize_predictions = _visualisation_only("visualize_predictions")
