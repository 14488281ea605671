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


class _MseFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, pred, target):
        if not pred.is_cuda:
            raise RuntimeError("CornerLoss: CUDA tensors required (no CPU fallback)")
        if pred.shape != target.shape:
            raise RuntimeError(f"CornerLoss: shape mismatch {tuple(pred.shape)} vs {tuple(target.shape)}")
        p = pred.detach().float().contiguous()
        t = target.detach().float().contiguous()
        lib = N.load()
        need_grad = pred.requires_grad
        dpred = torch.empty_like(p) if need_grad else None
        loss = torch.empty((), dtype=torch.float32, device=p.device)
        scratch = torch.empty(lib.mtgseg_mse_scratch_floats(), dtype=torch.float32, device=p.device)
        with torch.cuda.device(p.device):
            N.check(lib.mtgseg_mse_loss(p.data_ptr(), t.data_ptr(), N.ptr(dpred), loss.data_ptr(), scratch.data_ptr(), p.numel(),
                                        N.stream_ptr()), "mtgseg_mse_loss")
        ctx.save_for_backward(dpred) if need_grad else None
        ctx.in_dtype = pred.dtype
        return loss

    @staticmethod
    def backward(ctx, grad_out):
        (dpred,) = ctx.saved_tensors
        return (dpred * grad_out).to(ctx.in_dtype), None


class CornerLoss(nn.Module):
    """``CornerLoss`` (train-pose-estimation_custom/metrics.py:105-136): heatmap MSE, value and gradient from one fused pass."""

    def __init__(self, image_size=(480, 640), heatmap_size=(160, 120)):
        super().__init__()
        self.image_size = image_size
        self.heatmap_size = heatmap_size

    def forward(self, predictions: torch.Tensor, targets: torch.Tensor) -> torch.Tensor:
        return _MseFn.apply(predictions, targets)


class CornerMetrics:
    """``CornerMetrics`` (train-pose-estimation_custom/metrics.py:8-100): same constructor, ``reset`` / ``update`` / ``compute`` and
    result keys.  ``update`` is one kernel (argmax of both heatmaps, pixel distance, counts) accumulating on the device; nothing
    is copied to the host until ``compute``.  The reference keeps every distance in a Python list; here the accumulator is
    {fp64 sum, n, n(<=3 px), n(<=6 px)}, which is all ``compute`` needs."""

    def __init__(self, image_size=(480, 640)):
        self.image_size = image_size
        self._acc = None
        self.reset()

    def reset(self):
        if self._acc is not None:
            self._acc.zero_()

    def update(self, predictions: torch.Tensor, targets: torch.Tensor):
        if not predictions.is_cuda:
            raise RuntimeError("CornerMetrics: CUDA tensors required (no CPU fallback)")
        p = predictions.detach().float().contiguous()
        t = targets.detach().float().contiguous()
        if p.shape != t.shape or p.dim() != 4:
            raise RuntimeError(f"CornerMetrics: expected two (B,K,H,W) heatmap tensors, got {tuple(p.shape)} and {tuple(t.shape)}")
        if self._acc is None or self._acc.device != p.device:
            self._acc = torch.zeros(4, dtype=torch.int64, device=p.device)  # 32 bytes: {double sum; uint64 n, n3, n6}
        B, K, H, W = p.shape
        with torch.cuda.device(p.device):
            N.check(N.load().mtgseg_corner_metrics(p.data_ptr(), t.data_ptr(), self._acc.data_ptr(), B, K, H, W, float(self.image_size[0]),
                                                   float(self.image_size[1]), N.stream_ptr()), "mtgseg_corner_metrics")

    def compute(self):
        if self._acc is None:
            return {"corner_acc_3px": 0.0, "corner_acc_6px": 0.0, "mean_corner_distance": 0.0}
        host = self._acc.cpu()
        total = float(host[:1].view(torch.float64)[0])
        n, n3, n6 = (int(v) for v in host[1:])
        if n == 0:
            return {"corner_acc_3px": 0.0, "corner_acc_6px": 0.0, "mean_corner_distance": 0.0}
        return {"corner_acc_3px": n3 / n * 100, "corner_acc_6px": n6 / n * 100, "mean_corner_distance": total / n}
