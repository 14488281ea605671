"""``train/model.py`` surface on top of the sm_100a kernels.

``CardSegmentationModel`` keeps the reference's constructor / ``forward`` signature (train/model.py:18-48,
79-89), module tree and 319-entry ``state_dict`` layout, so ``load_state_dict``,
``torch.optim.AdamW``, ``torch.nn.utils.prune`` and the checkpoint helpers interoperate unchanged.  The
modules below only *hold parameters*: the arithmetic of a CUDA forward runs in ``libmtgseg_b200.so``.
Their ``forward`` methods are the ATen composite used exclusively by ``torch.onnx.export`` /
``torch.jit.trace`` (train/export.py:68-79,177-182), which must see standard ops to emit the 66-Conv
graph; a normal call never reaches them.
"""
from __future__ import annotations

from collections import OrderedDict

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import arch
from .engine import SegEngine


def _is_exporting() -> bool:
    return torch.jit.is_tracing() or torch.onnx.is_in_onnx_export()


class _ConvBNAct(nn.Sequential):
    """Conv2d(bias=False) -> BatchNorm2d -> activation, children named '0','1','2' (tv:ops/misc.py:69-118)."""

    def __init__(self, cin, cout, k=1, stride=1, dilation=1, groups=1, act=None):
        layers = [nn.Conv2d(cin, cout, k, stride, (k - 1) // 2 * dilation, dilation, groups, bias=False),
                  nn.BatchNorm2d(cout, eps=arch.BACKBONE_BN_EPS, momentum=arch.BACKBONE_BN_MOMENTUM)]
        if act == "HS":
            layers.append(nn.Hardswish(inplace=True))
        elif act == "RE":
            layers.append(nn.ReLU(inplace=True))
        super().__init__(*layers)


class _SqueezeExcite(nn.Module):
    """fc1/fc2 are biased 1x1 convs; ReLU then Hardsigmoid gate (tv:ops/misc.py:225-261)."""

    def __init__(self, channels, squeeze):
        super().__init__()
        self.avgpool = nn.AdaptiveAvgPool2d(1)
        self.fc1 = nn.Conv2d(channels, squeeze, 1)
        self.fc2 = nn.Conv2d(squeeze, channels, 1)
        self.activation = nn.ReLU()
        self.scale_activation = nn.Hardsigmoid()

    def forward(self, x):
        s = self.scale_activation(self.fc2(self.activation(self.fc1(self.avgpool(x)))))
        return s * x


class _InvertedResidual(nn.Module):
    """[1x1 expand] -> depthwise -> [SE] -> 1x1 project (+ input) (tv:models/mobilenetv3.py:53-115)."""

    def __init__(self, b: arch.Block):
        super().__init__()
        self.use_res_connect = b.stride == 1 and b.cin == b.cout
        layers = []
        if b.cexp != b.cin:
            layers.append(_ConvBNAct(b.cin, b.cexp, 1, act=b.act))
        layers.append(_ConvBNAct(b.cexp, b.cexp, b.kernel, 1 if b.dilation > 1 else b.stride, b.dilation, b.cexp, b.act))
        if b.use_se:
            layers.append(_SqueezeExcite(b.cexp, arch.make_divisible(b.cexp // 4, 8)))
        layers.append(_ConvBNAct(b.cexp, b.cout, 1, act=None))
        self.block = nn.Sequential(*layers)
        self.out_channels = b.cout

    def forward(self, x):
        y = self.block(x)
        return y + x if self.use_res_connect else y


class _Backbone(nn.ModuleDict):
    """features[0..16] keyed '0'..'16' like torchvision's IntermediateLayerGetter (tv:models/_utils.py:55-73)."""

    def __init__(self):
        layers = OrderedDict()
        layers["0"] = _ConvBNAct(3, arch.STEM_CHANNELS, 3, stride=2, act="HS")
        for i, b in enumerate(arch.BLOCKS, start=1):
            layers[str(i)] = _InvertedResidual(b)
        layers[str(arch.HIGH_FEATURE)] = _ConvBNAct(arch.BLOCKS[-1].cout, arch.HIGH_CHANNELS, 1, act="HS")
        super().__init__(layers)
        for m in self.modules():  # tv:models/mobilenetv3.py:198-208
            if isinstance(m, nn.Conv2d):
                nn.init.kaiming_normal_(m.weight, mode="fan_out")
                if m.bias is not None:
                    nn.init.zeros_(m.bias)
            elif isinstance(m, nn.BatchNorm2d):
                nn.init.ones_(m.weight)
                nn.init.zeros_(m.bias)

    def forward(self, x):
        out = OrderedDict()
        for name, module in self.items():
            x = module(x)
            if name == str(arch.LOW_FEATURE):
                out["low"] = x
        out["high"] = x
        return out


class LRASPPHead(nn.Module):
    """The reference's custom head (train/model.py:92-142): 3x3 ``cbr``, GAP->1x1->sigmoid ``scale``,
    x2 bilinear, biased 1x1 ``low_classifier`` + ``high_classifier``."""

    def __init__(self, high_channels, low_channels, num_classes, inter_channels=128):
        super().__init__()
        self.cbr = nn.Sequential(nn.Conv2d(high_channels, inter_channels, 3, padding=1, bias=False),
                                 nn.BatchNorm2d(inter_channels), nn.ReLU(inplace=True))
        self.scale = nn.Sequential(nn.AdaptiveAvgPool2d(1), nn.Conv2d(high_channels, inter_channels, 1, bias=False),
                                   nn.Sigmoid())
        self.low_classifier = nn.Conv2d(low_channels, num_classes, 1)
        self.high_classifier = nn.Conv2d(inter_channels, num_classes, 1)

    def forward(self, input):
        low, high = input["low"], input["high"]
        x = self.cbr(high) * self.scale(high)
        x = F.interpolate(x, size=low.shape[-2:], mode="bilinear", align_corners=False)
        return self.low_classifier(low) + self.high_classifier(x)


class _LRASPP(nn.Module):
    """backbone + classifier + final bilinear to the input size (tv:models/segmentation/lraspp.py:43-51)."""

    def __init__(self, num_classes, inter_channels):
        super().__init__()
        self.backbone = _Backbone()
        self.classifier = LRASPPHead(arch.HIGH_CHANNELS, arch.LOW_CHANNELS, num_classes, inter_channels)

    def forward(self, x):
        out = self.classifier(self.backbone(x))
        out = F.interpolate(out, size=x.shape[-2:], mode="bilinear", align_corners=False)
        return OrderedDict(out=out)


class _TrainStep(torch.autograd.Function):
    """Whole-network autograd node: forward = mtgseg_forward_train, backward = mtgseg_backward.  The parameters are
    passed as inputs only so that autograd routes their gradients; the kernels read them through the engine."""

    @staticmethod
    def forward(ctx, model, x, out_dtype, *params):
        tensors = model._step_tensors  # resolved by the caller: parameters, or masked non-leaf weights under active pruning
        logits, x32 = model.engine().train_forward(tensors, x, out_dtype)
        ctx.model, ctx.tensors, ctx.x32 = model, tensors, x32
        ctx.token = model.engine().train_token()
        ctx.param_ids = {id(p): i for i, p in enumerate(params)}
        return logits

    @staticmethod
    def backward(ctx, dlogits):
        model, tensors = ctx.model, ctx.tensors
        is_param = [id(t) in ctx.param_ids for t in tensors]
        flat, views = model.engine().train_backward(tensors, is_param, ctx.x32, dlogits, ctx.token, dp=model.data_parallel)
        model.last_flat_grad = flat  # one contiguous buffer: a data-parallel driver all-reduces it in one call
        grads = [None] * len(ctx.param_ids)
        for t, v in zip(tensors, views):
            if v is not None:
                grads[ctx.param_ids[id(t)]] = v
        return (None, None, None, *grads)


class CardSegmentationModel(nn.Module):
    """LR-ASPP / MobileNetV3-Large card segmenter (background 0, card 1) running on hand-written B200 kernels."""

    def __init__(self, num_classes=2, pretrained=True):
        super().__init__()
        if pretrained:
            raise RuntimeError(
                "pretrained=True needs torchvision's COCO checkpoint download (train/model.py:31-33); this build is "
                "offline. Use pretrained=False (train/config.py:23) and load_state_dict().")
        self.num_classes = num_classes
        self.model = _LRASPP(num_classes, inter_channels=128)
        self._engine = None
        self._state_cache = None
        self._state_sig = None
        self._has_masks = False
        self._slots = None
        self._param_slots = None
        # Arithmetic of an eval-mode forward.  "auto" follows the reference's own dtype rule: under torch.autocast
        # (train/train.py:96,142) the tensor-core path (bf16 storage, fp32 accumulate); without autocast (train/evaluate.py:66)
        # IEEE float32 end to end, within 1e-4 of the reference.  "bf16" / "fp32" force one of them.
        self.inference_precision = "auto"
        self.data_parallel = False  # parallel.enable_gradient_exchange: backward averages the gradients over the ranks itself
        self._ref_keys = list(self.state_dict().keys())  # the reference's 319-key layout (train/utils.py:227-280 checkpoints)
        self.last_flat_grad = None

    # -- engine plumbing ---------------------------------------------------------------------
    def engine(self) -> SegEngine:
        if self._engine is None:
            self._engine = SegEngine(self.num_classes, self.model.classifier.cbr[0].out_channels)
        return self._engine

    def _state_slots(self):
        """(module, attribute) of the 319 reference state entries, in reference order.  Captured once: the key list of a freshly
        built model IS the reference layout; torch.nn.utils.prune later renames `weight` -> `weight_orig` + `weight_mask` in the
        state_dict, but the effective tensor is still reachable as `module.weight`."""
        if self._slots is None:
            slots = []
            for key in self._ref_keys:
                prefix, _, attr = key.rpartition(".")
                slots.append((self.get_submodule(prefix) if prefix else self, attr))
            self._slots = slots
            # the slots that are Parameters in the reference layout: their identity validates the cache
            # (a model that was pruned BEFORE its first forward holds `<attr>_orig` instead: still a parameter slot)
            self._param_slots = [(m, a) for m, a in slots if a in m._parameters or (a + "_orig") in m._parameters]
        return self._slots

    def _state_tensors(self):
        """The 319 state tensors in reference order.  Cached (resolving them costs ~0.1 ms) and self-validating: the cache is
        keyed on the identity of the parameters, so module surgery that swaps Parameter objects (torch.nn.utils.prune apply /
        remove, train/prune.py:60-113) is picked up without any call from the user.  While pruning masks are active
        (evaluation and fine-tuning of a masked model, train/prune.py:144-175) the effective `weight = weight_orig * weight_mask` is
        refreshed through the pruning hook on every call (the native path never runs the child modules' forward, which is what
        normally triggers that hook) and nothing is cached; under autograd it is a non-leaf tensor, so gradients reach `weight_orig`
        already multiplied by the mask."""
        self._state_slots()
        # ~20 us: one dict lookup per parameter slot (walking self.parameters() costs 0.5 ms).  A pruned `weight` leaves
        # module._parameters (-> None), prune.remove registers a new Parameter object: both change the signature.
        sig = tuple(id(m._parameters.get(a)) for m, a in self._param_slots)
        if self._state_cache is not None and sig == self._state_sig and not self._has_masks:
            return self._state_cache
        self._has_masks = False
        for mod in {m for m, _ in self._state_slots()}:
            for hook in mod._forward_pre_hooks.values():
                if isinstance(hook, torch.nn.utils.prune.BasePruningMethod):
                    hook(mod, None)  # module.<name> = <name>_orig * <name>_mask
                    self._has_masks = True
        tensors = [getattr(m, a) for m, a in self._state_slots()]
        if self._has_masks and self._engine is not None:
            self._engine._stats_dirty = True  # recomputed tensors may reuse an address with version 0: always repack
        self._state_cache, self._state_sig = tensors, sig
        return tensors

    def invalidate_cache(self):
        self._state_cache = None

    def _apply(self, fn, *args, **kwargs):
        self._state_cache = None
        return super()._apply(fn, *args, **kwargs)

    def load_state_dict(self, *args, **kwargs):
        self._state_cache = None
        return super().load_state_dict(*args, **kwargs)

    def forward(self, x):
        """x float32 (B,3,H,W) -> logits (B,num_classes,H,W) (train/model.py:79-89)."""
        if _is_exporting():
            return self.model(x)["out"]
        if not x.is_cuda:
            raise RuntimeError("CardSegmentationModel runs on CUDA (sm_100a) only; got a CPU tensor and there is no "
                               "CPU fallback. Move the model and the batch to 'cuda'.")
        out_dtype = torch.get_autocast_dtype("cuda") if torch.is_autocast_enabled("cuda") else torch.float32
        if self.training:
            tensors = self._state_tensors()
            if torch.is_grad_enabled():
                req = [t for t in tensors if t.requires_grad]  # parameters, or masked non-leaf weights under active pruning
                if req:
                    if len(req) != len(self._param_slots):
                        raise RuntimeError("the CUDA training step computes all parameter gradients; freezing a subset is not supported")
                    self._step_tensors = tensors
                    return _TrainStep.apply(self, x, out_dtype, *req)
            return self.engine().train_forward(tensors, x, out_dtype)[0]
        return self.engine().infer(self._state_tensors(), x, logits_dtype=out_dtype, precision=self._precision(x))

    def _precision(self, x, override=None):
        mode = override or self.inference_precision
        if mode == "auto":
            return "bf16" if (torch.is_autocast_enabled("cuda") or x.dtype != torch.float32) else "fp32"
        if mode not in ("bf16", "fp32"):
            raise RuntimeError(f"inference_precision must be 'auto', 'bf16' or 'fp32', got {mode!r}")
        return mode

    @torch.no_grad()
    def predict(self, x, targets=None, want_logits=False, precision=None):
        """Batched inference as train/evaluate.py:66-78 uses it, fused: returns the uint8 argmax mask and, when
        ``targets`` is given, the int64[4] confusion counts; logits only on request.  ``precision``: None = the model's
        ``inference_precision`` rule (fp32-exact for a float32 batch outside autocast, like evaluate.py:66), or "bf16" / "fp32"."""
        return self.engine().infer(self._state_tensors(), x, logits_dtype=torch.float32 if want_logits else None,
                                   want_mask=True, targets=targets, precision=self._precision(x, precision))


def create_model(num_classes=2, pretrained=True):
    """train/model.py:145-156."""
    return CardSegmentationModel(num_classes=num_classes, pretrained=pretrained)


def count_parameters(model):
    """(total, trainable) parameter counts (train/model.py:159-172)."""
    total = sum(p.numel() for p in model.parameters())
    return total, sum(p.numel() for p in model.parameters() if p.requires_grad)


def get_model_size(model):
    """Parameters + buffers in MiB (train/model.py:175-193)."""
    nbytes = sum(p.nelement() * p.element_size() for p in model.parameters())
    nbytes += sum(b.nelement() * b.element_size() for b in model.buffers())
    return nbytes / 1024 / 1024
