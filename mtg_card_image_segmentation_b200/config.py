"""Hyper-parameters of the hot path; same names and values as the reference's ``train/config.py:8-71``
so that ``train.py`` / ``evaluate.py`` style drivers read them unchanged."""
import os

import torch


class Config:
    _HERE = os.path.dirname(os.path.abspath(__file__))
    DATASET_ROOT = os.path.realpath(os.path.join(_HERE, "..", "dataset"))
    TRAIN_IMAGE_DIR = os.path.join(DATASET_ROOT, "train", "images")
    TRAIN_MASK_DIR = os.path.join(DATASET_ROOT, "train", "masks")
    TEST_IMAGE_DIR = os.path.join(DATASET_ROOT, "test", "images")
    TEST_MASK_DIR = os.path.join(DATASET_ROOT, "test", "masks")

    MODEL_NAME = "lraspp_mobilenet_v3_large"
    NUM_CLASSES = 2
    INPUT_HEIGHT = 320
    INPUT_WIDTH = 240
    PRETRAINED = False

    BATCH_SIZE = 32
    NUM_EPOCHS = 100
    LEARNING_RATE = 1e-3
    WEIGHT_DECAY = 1e-4
    USE_AMP = True
    DICE_WEIGHT = 0.5
    BCE_WEIGHT = 0.5
    OPTIMIZER = "adamw"
    SCHEDULER = "cosine"
    WARMUP_EPOCHS = 5

    USE_AUGMENTATION = True
    ROTATION_LIMIT = 15
    BRIGHTNESS_LIMIT = 0.2
    CONTRAST_LIMIT = 0.2
    SATURATION_LIMIT = 0.2
    HUE_LIMIT = 0.1

    PATIENCE = 15
    SAVE_EVERY = 10
    VALIDATE_EVERY = 1
    CHECKPOINT_DIR = os.path.join(_HERE, "checkpoints")
    LOG_DIR = os.path.join(_HERE, "logs")

    DEVICE = torch.device("cuda" if torch.cuda.is_available() else "cpu")
    NUM_WORKERS = 4
    PIN_MEMORY = True
    METRICS = ["iou", "dice", "pixel_accuracy"]

    PRUNING_AMOUNT = 0.3
    PRUNING_STRUCTURED = False
    PRUNING_FINE_TUNE_EPOCHS = 20

    @classmethod
    def create_directories(cls):
        os.makedirs(cls.CHECKPOINT_DIR, exist_ok=True)
        os.makedirs(cls.LOG_DIR, exist_ok=True)

    @classmethod
    def print_config(cls):
        rows = [("Model", cls.MODEL_NAME), ("Input Size", f"{cls.INPUT_HEIGHT}x{cls.INPUT_WIDTH}"),
                ("Batch Size", cls.BATCH_SIZE), ("Learning Rate", cls.LEARNING_RATE), ("Epochs", cls.NUM_EPOCHS),
                ("Device", cls.DEVICE), ("Mixed Precision", cls.USE_AMP)]
        print("=" * 50)
        for k, v in rows:
            print(f"{k}: {v}")
        print("=" * 50)
