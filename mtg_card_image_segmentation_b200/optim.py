"""Fused multi-tensor AdamW (one kernel launch for all 178 parameter tensors) with torch.optim.AdamW's maths and
state_dict layout ('step', 'exp_avg', 'exp_avg_sq' per parameter), so checkpoints written by the reference's
``save_checkpoint`` (train/utils.py:243-249) load here and vice versa.  Stands in for ``create_optimizer``'s
``optim.AdamW(model.parameters(), lr, weight_decay)`` (train/train.py:167-171)."""
from __future__ import annotations

import numpy as np
import torch

from . import _native as N

_CHUNK = 1 << 14
_REC = np.dtype([("p", "<u8"), ("g", "<u8"), ("m", "<u8"), ("v", "<u8"), ("n", "<i4"), ("pad", "<i4")])


class FusedAdamW(torch.optim.Optimizer):
    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=1e-2):
        defaults = dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay, amsgrad=False, maximize=False,
                        foreach=None, capturable=False, differentiable=False, fused=None, decoupled_weight_decay=True)
        super().__init__(params, defaults)
        self._lib = N.load()
        self._tables = {}

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        for gi, group in enumerate(self.param_groups):
            params = [p for p in group["params"] if p.grad is not None]
            if not params:
                continue
            dev = params[0].device
            for p in params:
                if not p.is_cuda or p.dtype != torch.float32 or p.grad.dtype != torch.float32 or not p.grad.is_contiguous():
                    raise RuntimeError("FusedAdamW needs contiguous fp32 CUDA parameters and gradients (no CPU fallback)")
                st = self.state[p]
                if len(st) == 0:
                    st["step"] = torch.tensor(0.0, dtype=torch.float32)
                    st["exp_avg"] = torch.zeros_like(p, memory_format=torch.preserve_format)
                    st["exp_avg_sq"] = torch.zeros_like(p, memory_format=torch.preserve_format)
                st["step"] += 1
            steps = {int(self.state[p]["step"]) for p in (params[0], params[-1])}
            if len(steps) != 1:
                raise RuntimeError("FusedAdamW: parameters of one group must share the step count")
            step_no = steps.pop()
            # the chunk table only depends on pointers; the caching allocator hands the same gradient block back every
            # step, so in steady state the device copy is reused
            key = (gi, tuple(p.grad.data_ptr() for p in params), tuple(p.data_ptr() for p in params))
            if key not in self._tables:  # (the flat gradient buffer alternates between two allocator blocks)
                if len(self._tables) >= 8:
                    self._tables.clear()
                recs = []
                for p in params:
                    st = self.state[p]
                    n, base = p.numel(), (p.data_ptr(), p.grad.data_ptr(), st["exp_avg"].data_ptr(), st["exp_avg_sq"].data_ptr())
                    for off in range(0, n, _CHUNK):
                        recs.append((base[0] + 4 * off, base[1] + 4 * off, base[2] + 4 * off, base[3] + 4 * off, min(_CHUNK, n - off), 0))
                host = torch.from_numpy(np.array(recs, dtype=_REC).view(np.uint8).copy()).pin_memory()
                self._tables[key] = (host.to(dev, non_blocking=True), len(recs), host)
            table, nrec, _ = self._tables[key]
            b1, b2 = group["betas"]
            with torch.cuda.device(dev):
                N.check(self._lib.mtgseg_adamw_step(table.data_ptr(), nrec, float(group["lr"]), float(b1), float(b2),
                                                    float(group["eps"]), float(group["weight_decay"]), step_no, None, None,
                                                    N.stream_ptr()), "mtgseg_adamw_step")
            # the update happened outside autograd's view: bump the version counters (what the engine's packed-weight
            # cache keys on) with one multi-tensor no-op
            torch._foreach_add_(params, 0.0)
        return loss
