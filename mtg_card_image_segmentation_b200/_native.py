"""ctypes binding of libmtgseg_b200.so (the C ABI in include/mtgseg_b200.h).

The library is the product: if it is missing or does not load, every compute entry point raises —
nothing falls back to PyTorch or the CPU.
"""
from __future__ import annotations

import ctypes as C
import os
import threading

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libmtgseg_b200.so")

LOGITS_NONE, LOGITS_F32, LOGITS_BF16, LOGITS_F16 = 0, 1, 2, 3
ACT_NONE, ACT_RELU, ACT_HSWISH, ACT_HSIGMOID, ACT_SIGMOID = 0, 1, 2, 3, 4


class NetDesc(C.Structure):
    _fields_ = [("in_h", C.c_int32), ("in_w", C.c_int32), ("num_classes", C.c_int32), ("inter_channels", C.c_int32)]


class LayerProf(C.Structure):
    _fields_ = [("name", C.c_char * 48), ("kernel", C.c_char * 24), ("ms", C.c_float), ("bytes", C.c_double),
                ("flops", C.c_double)]


class PoseDesc(C.Structure):
    _fields_ = [("in_channels", C.c_int32), ("feat_h", C.c_int32), ("feat_w", C.c_int32), ("num_keypoints", C.c_int32),
                ("out_h", C.c_int32), ("out_w", C.c_int32)]


_vp, _i, _f, _sz = C.c_void_p, C.c_int, C.c_void_p, C.c_size_t
_ND = C.POINTER(NetDesc)

# name -> (restype, argtypes); mirrors include/mtgseg_b200.h one to one (tests/test_abi.py checks it)
SIGNATURES = {
    "mtgseg_version": (_i, []),
    "mtgseg_last_error": (C.c_char_p, []),
    "mtgseg_param_count": (_i, []),
    "mtgseg_packed_bytes": (_sz, [_ND]),
    "mtgseg_workspace_bytes": (_sz, [_ND, _i]),
    "mtgseg_pack_weights": (_i, [_ND, C.POINTER(_vp), _i, _vp, _vp]),
    "mtgseg_forward_infer": (_i, [_ND, _vp, _vp, _vp, _i, _vp, _vp, _vp, _vp, _sz, _i, _vp]),
    "mtgseg_workspace_bytes_f32": (_sz, [_ND, _i]),
    "mtgseg_forward_infer_f32": (_i, [_ND, _vp, C.POINTER(_vp), _i, _vp, _i, _vp, _vp, _vp, _vp, _sz, _i, _vp]),
    "mtgseg_train_workspace_bytes": (_sz, [_ND, _i]),
    "mtgseg_train_loss": (_i, [_ND, _vp, _vp, C.c_float, C.c_float, C.c_float, _vp, _sz, _i, _vp]),
    "mtgseg_forward_train": (_i, [_ND, _vp, _vp, C.POINTER(_vp), _i, _vp, _i, _vp, _sz, _i, _vp]),
    "mtgseg_backward": (_i, [_ND, _vp, _vp, C.POINTER(_vp), C.POINTER(_vp), _i, _vp, _i, _vp, _sz, _i, _vp, _sz, _i, _vp]),
    "mtgseg_dp_unique_id": (_i, [_vp]),
    "mtgseg_dp_init": (_i, [_vp, _i, _i]),
    "mtgseg_dp_world": (_i, []),
    "mtgseg_dp_shutdown": (_i, []),
    "mtgseg_dp_allreduce_avg": (_i, [_vp, _sz, _vp]),
    "mtgseg_adamw_step": (_i, [_vp, _i, C.c_float, C.c_float, C.c_float, C.c_float, C.c_float, _i, _vp, _vp, _vp]),
    "mtgseg_adamw_hyper": (_i, [_vp, C.c_float, C.c_float, C.c_float, C.c_float, C.c_float, _i, _vp]),
    "mtgseg_adamw_step_dev": (_i, [_vp, _i, _vp, _vp]),
    "mtgseg_bn_scratch_floats": (_sz, [_i, _i, _i]),
    "mtgseg_bn_train_fwd": (_i, [_vp, _vp, _vp, _vp, _vp, C.c_float, C.c_float, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i,
                                 _i, _i, _i, _i, _vp]),
    "mtgseg_bn_train_bwd": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _vp]),
    "mtgseg_wgrad": (_i, [_vp, _vp, _vp, _vp, _i, C.c_int64, _i, _i, _i, _i, _i, _vp]),
    "mtgseg_wgrad_tc": (_i, [_vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _i, _vp]),
    "mtgseg_dw_bwd": (_i, [_vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _i, _vp]),
    "mtgseg_stem_wgrad": (_i, [_vp, _vp, _vp, _i, _i, _i, _vp]),
    "mtgseg_upsample_bwd": (_i, [_vp, _i, _vp, _i, _i, _i, _i, _i, _i, _vp]),
    "mtgseg_se_bwd_scratch_floats": (_sz, [_i, _i, _i]),
    "mtgseg_se_block_bwd": (_i, [_vp, _vp, _vp, _vp, _vp, _i, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _vp]),
    "mtgseg_head_bwd_segments": (_i, [_i]),
    "mtgseg_head_bwd": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _i, _i, _vp]),
    "mtgseg_pose_param_count": (_i, []),
    "mtgseg_pose_packed_bytes": (_sz, [C.POINTER(PoseDesc)]),
    "mtgseg_pose_workspace_bytes": (_sz, [C.POINTER(PoseDesc), _i]),
    "mtgseg_pose_pack_weights": (_i, [C.POINTER(PoseDesc), C.POINTER(_vp), _i, _vp, _vp]),
    "mtgseg_pose_forward": (_i, [C.POINTER(PoseDesc), _vp, _vp, _vp, _vp, _vp, _sz, _i, _vp]),
    "mtgseg_decode_heatmaps": (_i, [_vp, _vp, _i, _i, _i, _i, _vp]),
    "mtgseg_corner_metrics": (_i, [_vp, _vp, _vp, _i, _i, _i, _i, C.c_float, C.c_float, _vp]),
    "mtgseg_mse_scratch_floats": (_sz, []),
    "mtgseg_mse_loss": (_i, [_vp, _vp, _vp, _vp, _vp, C.c_longlong, _vp]),
    "mtgseg_launch_count": (C.c_ulonglong, []),
    "mtgseg_forward_infer_profiled": (_i, [_ND, _vp, _vp, _vp, _i, _vp, _vp, _vp, _vp, _sz, _i, _vp, C.POINTER(LayerProf), _i,
                                           C.POINTER(_i)]),
    "mtgseg_forward_infer_u8": (_i, [_ND, _vp, _vp, _vp, _i, _vp, _vp, _vp, _vp, _sz, _i, _vp]),
    "mtgseg_metric_counts": (_i, [_vp, _i, _vp, _vp, C.c_int64, C.c_int64, _vp]),
    "mtgseg_loss_scratch_bytes": (_sz, []),
    "mtgseg_loss_fwd_bwd": (_i, [_vp, _i, _vp, _vp, _i, _vp, _vp, C.c_int64, C.c_int64, _i, C.c_float, C.c_float, C.c_float, _vp]),
    "mtgseg_loss_lowres_scratch_floats": (_sz, [_i, _i, _i]),
    "mtgseg_loss_lowres": (_i, [_vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _i, C.c_float, C.c_float, C.c_float, _vp]),
    "mtgseg_conv1x1": (_i, [_vp, _vp, _vp, _i, _i, _i, _vp, _vp, _i, _vp, _vp, _i, _vp]),
    "mtgseg_conv3x3": (_i, [_vp, _vp, _vp, _i, _i, _i, _i, _i, _vp, _vp, _i, _vp]),
    "mtgseg_dwconv_chunks": (_i, [_i, _i, _i, _i, _i, _i, _i]),
    "mtgseg_dwconv": (_i, [_vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _i, _vp, _vp, _i, _vp, _i, _vp]),
    "mtgseg_stem": (_i, [_vp, _vp, _vp, _vp, _vp, _i, _i, _i, _vp]),
    "mtgseg_se_mlp": (_i, [_vp, _i, _i, _i, _i, _i, _vp, _vp, _i, _vp, _vp, _i, _vp, _vp, _vp]),
    "mtgseg_gap": (_i, [_vp, _vp, _i, _i, _i, _vp]),
    "mtgseg_head_mix": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _i, _i, _vp]),
    "mtgseg_upsample_out": (_i, [_vp, _vp, _i, _vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _vp]),
}

_lock = threading.Lock()
_lib = None


def load() -> C.CDLL:
    """Load (once) and type the shared library.  Raises RuntimeError if it is not there."""
    global _lib
    if _lib is not None:
        return _lib
    with _lock:
        if _lib is None:
            if not os.path.exists(LIB_PATH):
                raise RuntimeError(
                    f"{LIB_PATH} is missing: build it with `python -m mtg_card_image_segmentation_b200.build` "
                    "(needs nvcc).  There is no fallback implementation.")
            lib = C.CDLL(LIB_PATH)
            for name, (res, args) in SIGNATURES.items():
                fn = getattr(lib, name)
                fn.restype, fn.argtypes = res, args
            _lib = lib
    return _lib


def check(rc: int, what: str) -> None:
    if rc != 0:
        msg = load().mtgseg_last_error()
        raise RuntimeError(f"{what} failed (code {rc}): {msg.decode() if msg else '?'}")


def ptr(t) -> int | None:
    """Device pointer of a torch tensor (None passes NULL)."""
    return None if t is None else t.data_ptr()


def stream_ptr() -> int:
    import torch
    return torch.cuda.current_stream().cuda_stream
