"""The fixed topology of the hot path: dilated MobileNetV3-Large feature extractor as torchvision's
``lraspp_mobilenet_v3_large`` builds it (tv:models/mobilenetv3.py:233-251 with ``dilated=True`` from
tv:models/segmentation/lraspp.py:172).  Must stay in sync with csrc/net.cu (tests/test_layout.py)."""
from collections import namedtuple

Block = namedtuple("Block", "cin kernel cexp cout use_se act stride dilation")

BLOCKS = (
    Block(16, 3, 16, 16, False, "RE", 1, 1),
    Block(16, 3, 64, 24, False, "RE", 2, 1),
    Block(24, 3, 72, 24, False, "RE", 1, 1),
    Block(24, 5, 72, 40, True, "RE", 2, 1),
    Block(40, 5, 120, 40, True, "RE", 1, 1),
    Block(40, 5, 120, 40, True, "RE", 1, 1),
    Block(40, 3, 240, 80, False, "HS", 2, 1),
    Block(80, 3, 200, 80, False, "HS", 1, 1),
    Block(80, 3, 184, 80, False, "HS", 1, 1),
    Block(80, 3, 184, 80, False, "HS", 1, 1),
    Block(80, 3, 480, 112, True, "HS", 1, 1),
    Block(112, 3, 672, 112, True, "HS", 1, 1),
    Block(112, 5, 672, 160, True, "HS", 2, 2),
    Block(160, 5, 960, 160, True, "HS", 1, 2),
    Block(160, 5, 960, 160, True, "HS", 1, 2),
)
STEM_CHANNELS = 16
LOW_FEATURE = 4       # backbone['4'] output -> 'low'  (40 channels, stride 8)
HIGH_FEATURE = 16     # backbone['16'] output -> 'high' (960 channels, stride 16)
LOW_CHANNELS = 40
HIGH_CHANNELS = 960
BACKBONE_BN_EPS, BACKBONE_BN_MOMENTUM = 1e-3, 1e-2  # tv:models/mobilenetv3.py:155


def make_divisible(v, divisor=8):
    new_v = max(divisor, int(v + divisor / 2) // divisor * divisor)
    if new_v < 0.9 * v:
        new_v += divisor
    return new_v
