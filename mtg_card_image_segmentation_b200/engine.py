"""Host-side driver of the inference path: owns the packed-weight arena and the activation workspace (torch
tensors, so PyTorch's caching allocator owns the memory) and calls the C ABI on torch's current stream."""
from __future__ import annotations

import ctypes as C

import torch

from . import _native as N

_TORCH_TO_LOGITS = {torch.float32: N.LOGITS_F32, torch.bfloat16: N.LOGITS_BF16, torch.float16: N.LOGITS_F16}


class SegEngine:
    def __init__(self, num_classes: int = 2, inter_channels: int = 128):
        self.lib = N.load()
        self.num_classes = num_classes
        self.inter_channels = inter_channels
        self._packed = None
        self._packed_sig = None
        self._ws = None
        self._ws_extra = {}
        self._train_ws = None
        self._static_flat = None
        self._stats_dirty = False
        self._desc_cache = {}
        self._train_gen = 0    # bumped by every training-mode forward (the ONE workspace holds the last forward's activations)
        self._pack_count = 0   # bumped by every re-pack of the weight arena

    def desc(self, h: int, w: int) -> N.NetDesc:
        key = (h, w)
        if key not in self._desc_cache:
            self._desc_cache[key] = N.NetDesc(h, w, self.num_classes, self.inter_channels)
        return self._desc_cache[key]

    # -- weights -----------------------------------------------------------------------------
    def pack(self, tensors, device) -> torch.Tensor:
        """(Re)pack the reference-layout state tensors when any of them changed (version counters)."""
        sig = (str(device), tuple((t.data_ptr(), t._version) for t in tensors))
        if self._packed is not None and sig == self._packed_sig and not self._stats_dirty:
            return self._packed
        self._stats_dirty = False
        n = self.lib.mtgseg_param_count()
        if len(tensors) != n:
            raise RuntimeError(f"expected {n} state_dict entries, got {len(tensors)}")
        for t in tensors:
            if t.device != device:
                raise RuntimeError("model parameters/buffers and the input batch must be on the same CUDA device")
            if not t.is_contiguous():
                raise RuntimeError("state tensors must be contiguous")
        d = self.desc(320, 240)  # packing does not depend on the input size
        nbytes = self.lib.mtgseg_packed_bytes(C.byref(d))
        if self._packed is None or self._packed.numel() != nbytes or self._packed.device != device:
            self._packed = torch.empty(nbytes, dtype=torch.uint8, device=device)
        arr = (C.c_void_p * n)(*[t.data_ptr() for t in tensors])
        N.check(self.lib.mtgseg_pack_weights(C.byref(d), arr, n, self._packed.data_ptr(), N.stream_ptr()), "mtgseg_pack_weights")
        self._packed_sig = sig
        self._pack_count += 1
        return self._packed

    def workspace(self, d: N.NetDesc, batch: int, device, slot=0) -> torch.Tensor:
        """Activation workspace; `slot` != 0 selects an independent buffer (concurrent sub-batches on other streams;
        ("f32", i) = the fp32-exact path's larger workspace)."""
        f32 = isinstance(slot, tuple)
        need = (self.lib.mtgseg_workspace_bytes_f32 if f32 else self.lib.mtgseg_workspace_bytes)(C.byref(d), batch)
        if need == 0:
            raise RuntimeError(f"mtgseg_workspace_bytes failed: {self.lib.mtgseg_last_error().decode()}")
        cur = self._ws if slot == 0 else self._ws_extra.get(slot)
        if cur is None or cur.numel() < need or cur.device != device:
            cur = torch.empty(need, dtype=torch.uint8, device=device)
            if slot == 0:
                self._ws = cur
            else:
                self._ws_extra[slot] = cur
        return cur

    # -- inference ---------------------------------------------------------------------------
    def infer(self, tensors, x, logits_dtype=torch.float32, want_mask=False, targets=None, ws_slot=0, out=None, mask_out=None,
              precision="bf16"):
        """`out` / `mask_out`: caller-owned logits / uint8 mask buffers (e.g. batch slices of a larger tensor).
        precision "bf16": tensor-core path (bf16 storage, fp32 accumulate); "fp32": IEEE fp32 end to end (train/evaluate.py:66)."""
        if precision not in ("bf16", "fp32"):
            raise RuntimeError(f"precision must be 'bf16' or 'fp32', got {precision!r}")
        u8 = x.dtype == torch.uint8
        if u8 and precision == "fp32":
            raise RuntimeError("the fp32-exact path takes the normalised float32 (B,3,H,W) batch, not raw uint8 frames")
        if u8:  # raw HWC pixels: normalisation is fused into the stem kernel
            if x.dim() != 4 or x.shape[3] != 3:
                raise RuntimeError(f"uint8 input must be (B,H,W,3) raw pixels, got {tuple(x.shape)}")
            x = x.contiguous()
            B, H, W, _ = x.shape
        else:
            if x.dim() != 4 or x.shape[1] != 3:
                raise RuntimeError(f"expected a (B,3,H,W) batch, got {tuple(x.shape)}")
            if x.dtype != torch.float32:
                x = x.float()
            x = x.contiguous()
            B, _, H, W = x.shape
        dev = x.device
        fwd = self.lib.mtgseg_forward_infer_u8 if u8 else self.lib.mtgseg_forward_infer
        with torch.cuda.device(dev):
            d = self.desc(H, W)
            if precision == "fp32":
                for t in tensors:
                    if t.device != dev or not t.is_contiguous():
                        raise RuntimeError("model parameters/buffers must be contiguous and on the input's CUDA device")
                ws = self.workspace(d, B, dev, ("f32", ws_slot))
            else:
                packed = self.pack(tensors, dev)
                ws = self.workspace(d, B, dev, ws_slot)
            logits = None
            if logits_dtype is not None:
                logits = out if out is not None else torch.empty((B, self.num_classes, H, W), dtype=logits_dtype, device=dev)
            mask = None
            if want_mask:
                mask = mask_out if mask_out is not None else torch.empty((B, H, W), dtype=torch.uint8, device=dev)
                if mask.dtype != torch.uint8 or tuple(mask.shape) != (B, H, W) or not mask.is_contiguous() or mask.device != dev:
                    raise RuntimeError("mask_out must be a contiguous uint8 (B,H,W) tensor on the input's device")
            counts = None
            if targets is not None:
                if targets.dtype != torch.int64 or tuple(targets.shape) != (B, H, W) or targets.device != dev:
                    raise RuntimeError("targets must be an int64 (B,H,W) tensor on the input's device")
                targets = targets.contiguous()
                counts = torch.zeros(4, dtype=torch.int64, device=dev)
            if precision == "fp32":
                rc = self.lib.mtgseg_forward_infer_f32(
                    C.byref(d), x.data_ptr(), self._ptr_array(tensors), len(tensors), N.ptr(logits),
                    _TORCH_TO_LOGITS.get(logits_dtype, N.LOGITS_NONE), N.ptr(mask), N.ptr(counts), N.ptr(targets),
                    ws.data_ptr(), ws.numel(), B, N.stream_ptr())
            else:
                rc = fwd(
                    C.byref(d), x.data_ptr(), packed.data_ptr(), N.ptr(logits),
                    _TORCH_TO_LOGITS.get(logits_dtype, N.LOGITS_NONE), N.ptr(mask), N.ptr(counts), N.ptr(targets),
                    ws.data_ptr(), ws.numel(), B, N.stream_ptr())
            N.check(rc, "mtgseg_forward_infer")
        if want_mask or targets is not None:
            return {"logits": logits, "mask": mask, "counts": counts}
        return logits


    # -- training ----------------------------------------------------------------------------
    def _ptr_array(self, tensors):
        return (C.c_void_p * len(tensors))(*[None if t is None else t.data_ptr() for t in tensors])

    def train_forward(self, tensors, x, logits_dtype=torch.float32):
        """model.train(); model(x): batch-statistics BatchNorm, running stats updated in place, activations saved.
        logits_dtype None: no full-resolution logits (the captured step takes loss and gradient from the low-resolution ones)."""
        if x.dim() != 4 or x.shape[1] != 3:
            raise RuntimeError(f"expected a (B,3,H,W) batch, got {tuple(x.shape)}")
        x = x.float().contiguous()
        B, _, H, W = x.shape
        dev = x.device
        with torch.cuda.device(dev):
            packed = self.pack(tensors, dev)
            d = self.desc(H, W)
            need = self.lib.mtgseg_train_workspace_bytes(C.byref(d), B)
            if need == 0:
                raise RuntimeError(f"mtgseg_train_workspace_bytes failed: {self.lib.mtgseg_last_error().decode()}")
            if self._train_ws is None or self._train_ws.numel() < need or self._train_ws.device != dev:
                self._train_ws = None
                self._train_ws = torch.empty(need, dtype=torch.uint8, device=dev)
            logits = None if logits_dtype is None else torch.empty((B, self.num_classes, H, W), dtype=logits_dtype, device=dev)
            rc = self.lib.mtgseg_forward_train(C.byref(d), x.data_ptr(), packed.data_ptr(), self._ptr_array(tensors), len(tensors),
                                               N.ptr(logits), _TORCH_TO_LOGITS.get(logits_dtype, N.LOGITS_NONE), self._train_ws.data_ptr(),
                                               self._train_ws.numel(), B, N.stream_ptr())
            N.check(rc, "mtgseg_forward_train")
        self._stats_dirty = True  # running statistics changed under the folded-BN cache
        self._train_gen += 1
        return logits, x

    def train_loss(self, x, targets, dice_weight, ce_weight, smooth, out=None):
        """CombinedLoss of the last training forward from its low-resolution logits (mtgseg_train_loss); leaves the pulled-back
        gradient in the workspace for train_backward(dlogits=None).  Returns loss3 = (total, dice, ce)."""
        B, _, H, W = x.shape
        t = targets.contiguous()
        if t.dtype != torch.int64 or tuple(t.shape) != (B, H, W) or t.device != x.device:
            raise RuntimeError("targets must be an int64 (B,H,W) tensor on the input's device")
        loss3 = out if out is not None else torch.empty(3, dtype=torch.float32, device=x.device)
        with torch.cuda.device(x.device):
            N.check(self.lib.mtgseg_train_loss(C.byref(self.desc(H, W)), t.data_ptr(), loss3.data_ptr(), dice_weight, ce_weight, smooth,
                                               self._train_ws.data_ptr(), self._train_ws.numel(), B, N.stream_ptr()), "mtgseg_train_loss")
        return loss3

    def train_token(self):
        """Identifies the forward whose activations the workspace holds and the weight arena it used (checked by backward)."""
        return (self._train_gen, self._pack_count)

    def train_backward(self, tensors, is_param, x, dlogits, token=None, dp=False):
        """loss.backward(): returns (flat fp32 gradient buffer, list of per-state-entry views or None).  `token` =
        train_token() taken right after the forward this backward belongs to."""
        if token is not None and token != self.train_token():
            what = ("another training-mode forward has overwritten the saved activations" if token[0] != self._train_gen else
                    "the weights were re-packed (optimizer step, load_state_dict or an eval-mode forward) after the forward")
            raise RuntimeError(f"backward() of a stale forward: {what}. The CUDA training step keeps ONE set of saved "
                               "activations per model: call backward() before the next model(x) / optimizer.step().")
        B, _, H, W = x.shape
        dev = x.device
        dlogits = None if dlogits is None else dlogits.contiguous()
        with torch.cuda.device(dev):
            sizes = [t.numel() if p else 0 for t, p in zip(tensors, is_param)]
            flat = self._static_flat  # GraphedTrainStep: one fixed gradient buffer, so that the captured addresses never change
            if flat is not None and flat.numel() == sum(sizes) and flat.device == dev:
                flat.zero_()
            else:
                flat = torch.zeros(sum(sizes), dtype=torch.float32, device=dev)
            views, off = [], 0
            for t, n in zip(tensors, sizes):
                views.append(flat[off:off + n].view_as(t) if n else None)
                off += n
            d = self.desc(H, W)
            rc = self.lib.mtgseg_backward(C.byref(d), x.data_ptr(), self._packed.data_ptr(), self._ptr_array(tensors),
                                          self._ptr_array(views), len(tensors), N.ptr(dlogits),
                                          N.LOGITS_NONE if dlogits is None else _TORCH_TO_LOGITS[dlogits.dtype],
                                          self._train_ws.data_ptr(), self._train_ws.numel(), B, flat.data_ptr(), flat.numel(),
                                          1 if dp else 0, N.stream_ptr())
            N.check(rc, "mtgseg_backward")
        return flat, views

    def profile(self, tensors, x, logits_dtype=torch.bfloat16):
        """One forward with CUDA events around every kernel launch -> list of dicts (bench.py roofline)."""
        x = x.contiguous()
        B, _, H, W = x.shape
        dev = x.device
        with torch.cuda.device(dev):
            packed = self.pack(tensors, dev)
            d = self.desc(H, W)
            ws = self.workspace(d, B, dev)
            logits = torch.empty((B, self.num_classes, H, W), dtype=logits_dtype, device=dev)
            recs = (N.LayerProf * 128)()
            n = C.c_int(0)
            rc = self.lib.mtgseg_forward_infer_profiled(
                C.byref(d), x.data_ptr(), packed.data_ptr(), logits.data_ptr(), _TORCH_TO_LOGITS[logits_dtype], None, None,
                None, ws.data_ptr(), ws.numel(), B, N.stream_ptr(), recs, 128, C.byref(n))
            N.check(rc, "mtgseg_forward_infer_profiled")
        return [{"name": recs[i].name.decode(), "kernel": recs[i].kernel.decode(), "ms": recs[i].ms,
                 "bytes": recs[i].bytes, "flops": recs[i].flops} for i in range(n.value)]


class GraphedInference:
    """CUDA-graph replay of one fixed-shape inference call (launch-bound regimes: ~60 kernels per forward).

    ``splits`` > 1 runs that many sub-batches concurrently on forked streams inside the graph (images are independent
    units): kernels of different sub-batches fill each other's ramp-up / drain phases.
    ``run(x)`` copies ``x`` into the captured input buffer and replays; outputs are the captured tensors."""

    def __init__(self, model, example, logits_dtype=torch.bfloat16, want_mask=False, splits=1, precision="bf16"):
        self.model = model
        self._precision = precision
        self.x = example.clone()
        eng, tensors = model.engine(), model._state_tensors()
        B = self.x.shape[0]
        splits = max(1, min(splits, B))
        bounds = [(i * B) // splits for i in range(splits + 1)]
        # spatial size of the input: (B,3,H,W) float batches or (B,H,W,3) uint8 frames; logits_dtype None = mask / counts only
        hw = tuple(self.x.shape[1:3]) if self.x.dtype == torch.uint8 else tuple(self.x.shape[2:])
        logits = None if logits_dtype is None else torch.empty((B, eng.num_classes) + hw, dtype=logits_dtype, device=self.x.device)

        mask = torch.empty((B,) + hw, dtype=torch.uint8, device=self.x.device) if want_mask and splits > 1 else None

        def run_all():
            if splits == 1:
                return eng.infer(tensors, self.x, logits_dtype=logits_dtype, want_mask=want_mask, out=logits, precision=precision)
            main = torch.cuda.current_stream()
            for i in range(splits):
                st = self._streams[i]
                st.wait_stream(main)
                with torch.cuda.stream(st):
                    sl = slice(bounds[i], bounds[i + 1])
                    eng.infer(tensors, self.x[sl], logits_dtype=logits_dtype, want_mask=want_mask, ws_slot=i,
                              out=None if logits is None else logits[sl], mask_out=None if mask is None else mask[sl],
                              precision=precision)
            for st in self._streams:
                main.wait_stream(st)
            return {"logits": logits, "mask": mask, "counts": None} if want_mask else logits

        self._streams = [torch.cuda.Stream() for _ in range(splits)] if splits > 1 else []
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):  # warm-up outside capture: packs weights, sizes the workspaces, sets func attributes
            run_all()
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        before = eng.lib.mtgseg_launch_count()
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self.out = run_all()
        self.launches_per_replay = int(eng.lib.mtgseg_launch_count() - before)

    def replay(self):
        # the captured kernels read the packed weight arena, not the parameters: after an optimizer step / load_state_dict the
        # arena is refreshed here, in place (same address, so the graph stays valid), on the current stream, before the replay.
        # Costs one signature comparison (~0.05 ms of host time) per replay when nothing changed.
        if self._precision == "bf16":
            self.model.engine().pack(self.model._state_tensors(), self.x.device)
        self.graph.replay()
        return self.out

    def run(self, x):
        self.x.copy_(x, non_blocking=True)
        return self.replay()


class GraphedTrainStep:
    """The body of train/train.py:88-110's loop -- ``optimizer.zero_grad(); loss = criterion(model(x), y); loss.backward();
    optimizer.step()`` -- captured once and replayed as ONE CUDA graph (weight re-pack, ~480 kernels on two streams, AdamW).

    At batch 32 the step is a chain of short dependent launches; the graph removes the per-launch gaps.  ``step(x, y)`` copies the
    batch into the captured buffers, refreshes the optimizer's device hyperparameter block (so LR schedulers keep working) and
    replays; it returns the captured loss scalar (a device tensor: read it when you need it).  Building the object leaves model,
    BatchNorm statistics and optimizer state exactly as they were (the warm-up steps run on a snapshot that is restored).
    Construct it under the same ``torch.autocast`` context as the loop.  Not supported: GradScaler (bf16 needs none), several
    parameter groups, criteria other than this package's CombinedLoss / DiceLoss (the step calls the engine and the fused loss
    kernel directly, without autograd), active pruning masks.  Data parallel: call ``parallel.enable_gradient_exchange(model)`` first;
    the bucketed NCCL exchange is then part of the captured backward."""

    def __init__(self, model, criterion, optimizer, example_x, example_y, warmup=3, allow_unsynchronised=False, lowres_loss=False):
        from .optim import FusedAdamW
        # lowres_loss: take loss and gradient from the head's 40x30 logits (mtgseg_train_loss): no full-resolution logits / dlogits
        # tensors in the step (-59 MB at B=32) and two launches fewer.  Measured on a B200: neutral at B=32 (5.43 vs 5.41 ms), slower
        # at B=256 (26.41 vs 26.04 ms: every fine pixel's softmax is recomputed by its four corner owners), hence off by default.
        self._lowres_loss = bool(lowres_loss)
        if not isinstance(optimizer, FusedAdamW) or len(optimizer.param_groups) != 1:
            raise RuntimeError("GraphedTrainStep needs a FusedAdamW optimizer with one parameter group")
        if torch.distributed.is_available() and torch.distributed.is_initialized() and torch.distributed.get_world_size() > 1 \
                and not getattr(model, "data_parallel", False) and not allow_unsynchronised:
            raise RuntimeError("GraphedTrainStep in a multi-process job needs parallel.enable_gradient_exchange(model) first: the "
                               "captured step contains the bucketed NCCL gradient exchange of mtgseg_backward")
        if not model.training:
            raise RuntimeError("GraphedTrainStep: call model.train() first")
        if not example_x.is_cuda:
            raise RuntimeError("GraphedTrainStep runs on CUDA (sm_100a) only")
        self.model, self.criterion, self.optimizer = model, criterion, optimizer
        dev = example_x.device
        eng = model.engine()
        tensors = model._state_tensors()
        if model._has_masks:
            raise RuntimeError("GraphedTrainStep does not support active pruning masks (the masked weights are rebuilt on the host)")
        self._params = list(optimizer.param_groups[0]["params"])
        self._is_param = [bool(t.requires_grad) for t in tensors]
        if sum(self._is_param) != len(self._params) or {id(p) for p in self._params} != {id(t) for t in tensors if t.requires_grad}:
            raise RuntimeError("GraphedTrainStep: the optimizer must hold exactly the model's parameters, all trainable")
        from .utils import loss_weights
        self._loss_weights = loss_weights(criterion)
        self._logits_dtype = torch.get_autocast_dtype("cuda") if torch.is_autocast_enabled("cuda") else torch.float32
        self.x = example_x.detach().float().contiguous().clone()
        self.y = example_y.detach().contiguous().clone()
        if self.x.dim() != 4 or self.x.shape[1] != 3 or self.y.dtype != torch.int64 or self.y.device != dev or \
                tuple(self.y.shape) != (self.x.shape[0],) + tuple(self.x.shape[2:]):
            raise RuntimeError("GraphedTrainStep: expected a (B,3,H,W) batch and int64 (B,H,W) targets on the same CUDA device")
        self.hyper = torch.zeros(8, dtype=torch.float32, device=dev)
        self._opt_epoch = getattr(optimizer, "_state_epoch", 0)
        with torch.no_grad():
            snap = [t.detach().clone() for t in tensors]
            snap_opt = []
            for p in self._params:
                st = optimizer.state.get(p)
                snap_opt.append((st["step"].clone(), st["exp_avg"].clone(), st["exp_avg_sq"].clone()) if st else None)
        eng._static_flat = torch.empty(sum(p.numel() for p in self._params), dtype=torch.float32, device=dev)
        try:
            cur = torch.cuda.current_stream(dev)
            side = torch.cuda.Stream(dev)
            side.wait_stream(cur)
            with torch.cuda.stream(side):
                self._forward_backward()
                optimizer.step()  # eager: creates the optimizer state and the chunk table for the static gradient buffer
                for _ in range(max(1, warmup - 1)):
                    self._one_step()
            cur.wait_stream(side)
            torch.cuda.synchronize(dev)
            eng._stats_dirty = True  # the capture must contain the weight re-pack
            before = eng.lib.mtgseg_launch_count()
            self.graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(self.graph):
                self.loss = self._captured_body()
            self.launches_per_replay = int(eng.lib.mtgseg_launch_count() - before)
        finally:
            self._keep = (eng._static_flat, eng._train_ws, eng._packed, getattr(self, "_table", None))  # addresses the graph holds
            eng._static_flat = None
        with torch.no_grad():  # undo the warm-up steps
            for t, s in zip(tensors, snap):
                t.copy_(s)
            for p, s in zip(self._params, snap_opt):
                st = optimizer.state[p]
                if s is None:
                    st["step"].zero_(); st["exp_avg"].zero_(); st["exp_avg_sq"].zero_()
                else:
                    st["step"].copy_(s[0]); st["exp_avg"].copy_(s[1]); st["exp_avg_sq"].copy_(s[2])
        eng._stats_dirty = True

    def _forward_backward(self):
        # straight through the engine, no autograd: the autograd engine would run AccumulateGrad nodes on whatever stream they were
        # created on (e.g. the default stream of earlier eager steps whose loss tensor is still alive), which a capture forbids
        model, eng = self.model, self.model.engine()
        tensors = model._state_tensors()
        from .utils import fused_loss
        dp = getattr(model, "data_parallel", False)
        if self._lowres_loss:
            _, x32 = eng.train_forward(tensors, self.x, None)
            loss3 = eng.train_loss(x32, self.y, *self._loss_weights)
            flat, views = eng.train_backward(tensors, self._is_param, x32, None, dp=dp)
        else:
            logits, x32 = eng.train_forward(tensors, self.x, self._logits_dtype)
            loss3, dlogits = fused_loss(logits, self.y, *self._loss_weights, True)
            flat, views = eng.train_backward(tensors, self._is_param, x32, dlogits, dp=dp)
        model.last_flat_grad = flat
        for t, v in zip(tensors, views):
            if v is not None:
                t.grad = v
        return loss3[0]

    def _captured_body(self):
        loss = self._forward_backward()
        self._table = self.optimizer.step_captured(self.hyper)
        return loss

    def _one_step(self):
        self.optimizer.advance(self.hyper)
        return self._captured_body()

    def step(self, x, y):
        if getattr(self.optimizer, "_state_epoch", 0) != self._opt_epoch:
            raise RuntimeError("optimizer.load_state_dict() replaced the moment tensors the captured AdamW launch updates: build a "
                               "new GraphedTrainStep after loading a checkpoint")
        self.x.copy_(x, non_blocking=True)
        self.y.copy_(y, non_blocking=True)
        self.optimizer.advance(self.hyper)
        self.optimizer._opt_called = True  # what torch's LR schedulers check before their own step()
        self.graph.replay()
        torch.autograd.graph.increment_version(self._params)  # parameters changed behind autograd's back
        self.model.engine()._stats_dirty = True               # so did the BatchNorm running statistics
        return self.loss
