"""Corner-keypoint head of the reference's pose pipeline on the sm_100a kernels.

Mirrors ``HRNetPoseHead`` (train-pose-estimation_custom/model.py:10-77: same constructor, module tree and 28-entry
``state_dict``) and ``LiteHRNet.decode_heatmaps`` (model.py:133-164).  The ``timm`` HRNet backbone of ``LiteHRNet``
(model.py:92-97) is third-party, needs a network download and is not importable offline: it is out of scope, the
head consumes the backbone's feature map.  CUDA only, eval mode only (the pose trainer is a
separate pipeline).
"""
from __future__ import annotations

import ctypes as C

import torch
import torch.nn as nn

from . import _native as N


class HRNetPoseHead(nn.Module):
    def __init__(self, in_channels: int, num_keypoints: int = 4, target_size=(160, 120)):
        super().__init__()
        self.num_keypoints = num_keypoints
        self.target_size = target_size  # (width, height)
        self.deconv_layers = nn.ModuleList([
            nn.Sequential(nn.ConvTranspose2d(in_channels, 256, kernel_size=4, stride=2, padding=1, bias=False),
                          nn.BatchNorm2d(256), nn.ReLU(inplace=True)),
            nn.Sequential(nn.ConvTranspose2d(256, 256, kernel_size=4, stride=2, padding=1, bias=False),
                          nn.BatchNorm2d(256), nn.ReLU(inplace=True)),
        ])
        self.conv_layers = nn.Sequential(
            nn.Conv2d(256, 256, kernel_size=3, padding=1), nn.BatchNorm2d(256), nn.ReLU(inplace=True),
            nn.Conv2d(256, 256, kernel_size=3, padding=1), nn.BatchNorm2d(256), nn.ReLU(inplace=True))
        self.final_layer = nn.Conv2d(256, num_keypoints, kernel_size=1, stride=1, padding=0)
        self.adaptive_pool = nn.AdaptiveAvgPool2d(target_size[::-1])  # (height, width)
        self._packed = None
        self._sig = None
        self._ws = None

    def _composite(self, x):  # ATen path for torch.onnx.export / torch.jit.trace only
        for d in self.deconv_layers:
            x = d(x)
        return self.adaptive_pool(self.final_layer(self.conv_layers(x)))

    def forward(self, x: torch.Tensor, return_coords: bool = False):
        """x: backbone features float32 (B, in_channels, h, w) -> heatmaps float32 (B, num_keypoints, H, W)."""
        if torch.jit.is_tracing() or torch.onnx.is_in_onnx_export():
            return self._composite(x)
        if not x.is_cuda:
            raise RuntimeError("HRNetPoseHead runs on CUDA (sm_100a) only; there is no CPU fallback")
        if self.training:
            raise NotImplementedError("the pose head kernels implement eval-mode inference (running-statistics BatchNorm)")
        lib = N.load()
        x = x.float().contiguous()
        B, Cin, Hf, Wf = x.shape
        d = N.PoseDesc(Cin, Hf, Wf, self.num_keypoints, self.target_size[1], self.target_size[0])
        tensors = list(self.state_dict(keep_vars=True).values())
        sig = tuple((t.data_ptr(), t._version) for t in tensors)
        with torch.cuda.device(x.device):
            if self._packed is None or sig != self._sig:
                nbytes = lib.mtgseg_pose_packed_bytes(C.byref(d))
                if nbytes == 0:
                    raise RuntimeError(lib.mtgseg_last_error().decode())
                self._packed = torch.empty(nbytes, dtype=torch.uint8, device=x.device)
                arr = (C.c_void_p * len(tensors))(*[t.data_ptr() for t in tensors])
                N.check(lib.mtgseg_pose_pack_weights(C.byref(d), arr, len(tensors), self._packed.data_ptr(), N.stream_ptr()),
                        "mtgseg_pose_pack_weights")
                self._sig = sig
            need = lib.mtgseg_pose_workspace_bytes(C.byref(d), B)
            if self._ws is None or self._ws.numel() < need or self._ws.device != x.device:
                self._ws = torch.empty(need, dtype=torch.uint8, device=x.device)
            hm = torch.empty((B, self.num_keypoints, d.out_h, d.out_w), dtype=torch.float32, device=x.device)
            coords = torch.empty((B, 2 * self.num_keypoints), dtype=torch.float32, device=x.device) if return_coords else None
            N.check(lib.mtgseg_pose_forward(C.byref(d), x.data_ptr(), self._packed.data_ptr(), hm.data_ptr(), N.ptr(coords),
                                            self._ws.data_ptr(), self._ws.numel(), B, N.stream_ptr()), "mtgseg_pose_forward")
        return (hm, coords) if return_coords else hm


def decode_heatmaps(heatmaps: torch.Tensor) -> torch.Tensor:
    """``LiteHRNet.decode_heatmaps`` (model.py:133-164): per-keypoint argmax -> (x, y) normalised to [0, 1], interleaved."""
    if not heatmaps.is_cuda:
        raise RuntimeError("decode_heatmaps: CUDA tensors required (no CPU fallback)")
    h = heatmaps.float().contiguous()
    B, K, H, W = h.shape
    coords = torch.empty((B, 2 * K), dtype=torch.float32, device=h.device)
    with torch.cuda.device(h.device):
        N.check(N.load().mtgseg_decode_heatmaps(h.data_ptr(), coords.data_ptr(), B, K, H, W, N.stream_ptr()), "mtgseg_decode_heatmaps")
    return coords
