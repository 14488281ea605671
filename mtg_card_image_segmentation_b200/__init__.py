"""B200-native (sm_100a) drop-in for the reference's segmentation hot path.

Mirrors the reference surface (``train/model.py``, ``train/utils.py``, ``train/config.py``):

    from mtg_card_image_segmentation_b200 import create_model, CombinedLoss, MetricsCalculator, Config

All device work goes through ``libmtgseg_b200.so`` (hand-written CUDA behind the C ABI of
``include/mtgseg_b200.h``); there is no CPU or eager-PyTorch fallback for the compute path.
"""
from .config import Config
from .model import CardSegmentationModel, LRASPPHead, count_parameters, create_model, get_model_size
from .utils import (CombinedLoss, DiceLoss, MetricsCalculator, calculate_dice_coefficient, calculate_iou,
                    calculate_pixel_accuracy, confusion_counts, load_checkpoint, save_checkpoint)

__all__ = [
    "Config", "CardSegmentationModel", "LRASPPHead", "create_model", "count_parameters", "get_model_size",
    "CombinedLoss", "DiceLoss", "MetricsCalculator", "calculate_iou", "calculate_dice_coefficient",
    "calculate_pixel_accuracy", "confusion_counts", "save_checkpoint", "load_checkpoint",
]
