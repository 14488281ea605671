"""``train/evaluate.py`` surface: ``ModelEvaluator.evaluate_dataset`` (evaluate.py:41-100) and
``_calculate_per_class_metrics`` (evaluate.py:102-137) on top of the fused inference call.

The reference pushes every predicted pixel through Python lists into sklearn's ``confusion_matrix`` (its real
bottleneck).  Here the argmax mask and the 2x2 integer confusion counts come out of the same kernel
launch as the logits, so the confusion matrix is exact int64 arithmetic on the device and no pixel ever visits the
host unless the caller asks for the masks (``keep_predictions=True`` returns uint8 arrays instead of Python lists).
Plots and failure mining (evaluate.py:139-330) are visualisation and stay with the reference.
"""
from __future__ import annotations

import numpy as np
import torch

from .utils import MetricsCalculator, per_class_metrics


class ModelEvaluator:
    def __init__(self, model, device, num_classes=2):
        self.model = model
        self.device = device
        self.num_classes = num_classes
        self.model.eval()

    def evaluate_dataset(self, dataloader, criterion=None, keep_predictions=False):
        """Same result dict as the reference: basic_metrics, confusion_matrix (np.int64 [target, prediction]),
        per_class_metrics, predictions, targets, filenames."""
        metrics_calc = MetricsCalculator(num_classes=self.num_classes, device=self.device)
        total = torch.zeros(4, dtype=torch.int64, device=self.device)
        preds, tgts, names = [], [], []
        with torch.no_grad():
            for batch in dataloader:
                images = batch["image"].to(self.device, non_blocking=True)
                masks = batch["mask"].to(self.device, non_blocking=True)
                out = self.model.predict(images, targets=masks, want_logits=criterion is not None)
                total += out["counts"]
                if criterion is not None:
                    loss = criterion(out["logits"], masks)
                    metrics_calc.update(loss, out["logits"], masks)
                if keep_predictions:
                    preds.append(out["mask"].cpu().numpy().reshape(-1))
                    tgts.append(masks.to(torch.uint8).cpu().numpy().reshape(-1))
                names.extend(batch.get("filename", []))
        cm = total.reshape(2, 2).cpu().numpy().astype(np.int64)
        return {
            "basic_metrics": metrics_calc.get_metrics(),
            "confusion_matrix": cm,
            "per_class_metrics": self._calculate_per_class_metrics(cm),
            "predictions": np.concatenate(preds) if preds else [],
            "targets": np.concatenate(tgts) if tgts else [],
            "filenames": names,
        }

    def _calculate_per_class_metrics(self, cm):
        return per_class_metrics(cm)
