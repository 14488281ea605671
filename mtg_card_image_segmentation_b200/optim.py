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

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        for group in self.param_groups:
            recs, step_no, dev = [], None, None
            for p in group["params"]:
                if p.grad is None:
                    continue
                if not p.is_cuda or p.dtype != torch.float32 or p.grad.dtype != torch.float32:
                    raise RuntimeError("FusedAdamW needs fp32 CUDA parameters and gradients (no CPU fallback)")
                st = self.state[p]
                if len(st) == 0:
                    st["step"] = torch.tensor(0.0, dtype=torch.float32)
                    st["exp_avg"] = torch.zeros_like(p, memory_format=torch.preserve_format)
                    st["exp_avg_sq"] = torch.zeros_like(p, memory_format=torch.preserve_format)
                st["step"] += 1
                s = int(st["step"].item()) if st["step"].device.type == "cpu" else int(st["step"])
                if step_no is None:
                    step_no, dev = s, p.device
                elif s != step_no:
                    raise RuntimeError("FusedAdamW: parameters of one group must share the step count")
                g = p.grad if p.grad.is_contiguous() else p.grad.contiguous()
                n, base = p.numel(), (p.data_ptr(), g.data_ptr(), st["exp_avg"].data_ptr(), st["exp_avg_sq"].data_ptr())
                for off in range(0, n, _CHUNK):
                    recs.append((base[0] + 4 * off, base[1] + 4 * off, base[2] + 4 * off, base[3] + 4 * off, min(_CHUNK, n - off), 0))
            if not recs:
                continue
            table = torch.from_numpy(np.array(recs, dtype=_REC).view(np.uint8).copy()).to(dev, non_blocking=True)
            b1, b2 = group["betas"]
            with torch.cuda.device(dev):
                N.check(self._lib.mtgseg_adamw_step(table.data_ptr(), len(recs), float(group["lr"]), float(b1), float(b2),
                                                    float(group["eps"]), float(group["weight_decay"]), step_no, None, None,
                                                    N.stream_ptr()), "mtgseg_adamw_step")
            self._keep = table  # keep the table alive until the kernel has run
            # the update happened outside autograd's view: bump the version counters (what the engine's packed-weight
            # cache keys on) with one multi-tensor no-op
            torch._foreach_add_([p for p in group["params"] if p.grad is not None], 0.0)
        return loss
