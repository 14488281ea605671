"""The C-ABI library loads and exports exactly the symbols include/mtgseg_b200.h declares (no compute calls:
these run without a GPU)."""
import ctypes
import os
import re

import pytest

from conftest import ROOT

HEADER = os.path.join(ROOT, "include", "mtgseg_b200.h")


def _declared():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(mtgseg_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    from mtg_card_image_segmentation_b200 import _native as N
    if not os.path.exists(N.LIB_PATH):
        from mtg_card_image_segmentation_b200 import build
        build.build()
    lib = ctypes.CDLL(N.LIB_PATH)
    names = _declared()
    assert len(names) >= 18
    for n in names:
        assert hasattr(lib, n), f"{n} declared in the header but not exported"


def test_binding_covers_header_and_basic_queries():
    from mtg_card_image_segmentation_b200 import _native as N
    assert sorted(N.SIGNATURES) == _declared()
    lib = N.load()
    assert lib.mtgseg_version() == 2
    assert lib.mtgseg_param_count() == 319
    d = N.NetDesc(320, 240, 2, 128)
    packed = lib.mtgseg_packed_bytes(ctypes.byref(d))
    assert 12_000_000 < packed < 19_000_000  # bf16 weights + their transposed dgrad copies + folded BN constants
    ws1 = lib.mtgseg_workspace_bytes(ctypes.byref(d), 1)
    ws8 = lib.mtgseg_workspace_bytes(ctypes.byref(d), 8)
    assert 16_000_000 < ws1 < 20_000_000 and 7.5 * ws1 < ws8 < 8.5 * ws1  # ~16.6 MB of bf16 activations per image
    assert lib.mtgseg_loss_scratch_bytes() > 0


def test_errors_are_reported_not_swallowed():
    from mtg_card_image_segmentation_b200 import _native as N
    lib = N.load()
    bad = N.NetDesc(320, 240, 99, 128)
    assert lib.mtgseg_packed_bytes(ctypes.byref(bad)) == 0
    assert b"num_classes" in lib.mtgseg_last_error()
    rc = lib.mtgseg_forward_infer(ctypes.byref(N.NetDesc(320, 240, 2, 128)), None, None, None, 1, None, None, None, None, 0, 1, None)
    assert rc == -1 and b"null" in lib.mtgseg_last_error()
    with pytest.raises(RuntimeError, match="null"):
        N.check(rc, "mtgseg_forward_infer")
