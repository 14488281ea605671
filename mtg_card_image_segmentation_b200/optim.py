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

    def _table(self, gi, params):
        """Device chunk table of one parameter group.  It only depends on pointers; the caching allocator hands the same gradient
        block back every step, so in steady state the device copy is reused."""
        dev = params[0].device
        # every pointer a record embeds is part of the key: Optimizer.load_state_dict (utils.load_checkpoint) REPLACES the
        # exp_avg / exp_avg_sq tensors, and a table built before the load would keep updating the freed ones
        key = (gi, tuple(p.grad.data_ptr() for p in params), tuple(p.data_ptr() for p in params),
               tuple(self.state[p]["exp_avg"].data_ptr() for p in params),
               tuple(self.state[p]["exp_avg_sq"].data_ptr() for p in params))
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
        return self._tables[key]

    def load_state_dict(self, state_dict):
        super().load_state_dict(state_dict)
        self._tables.clear()  # the moment tensors were replaced
        self._state_epoch = getattr(self, "_state_epoch", 0) + 1  # engine.GraphedTrainStep refuses to replay a stale capture

    def add_param_group(self, param_group):
        super().add_param_group(param_group)
        if hasattr(self, "_tables"):
            self._tables.clear()

    def _checked_params(self, group, init_state):
        params = [p for p in group["params"] if p.grad is not None]
        for p in params:
            if not p.is_cuda or p.dtype != torch.float32 or p.grad.dtype != torch.float32 or not p.grad.is_contiguous():
                raise RuntimeError("FusedAdamW needs contiguous fp32 CUDA parameters and gradients (no CPU fallback)")
            st = self.state[p]
            if len(st) == 0:
                if not init_state:
                    raise RuntimeError("FusedAdamW: optimizer state must exist before a step is captured")
                st["step"] = torch.tensor(0.0, dtype=torch.float32)
                st["exp_avg"] = torch.zeros_like(p, memory_format=torch.preserve_format)
                st["exp_avg_sq"] = torch.zeros_like(p, memory_format=torch.preserve_format)
        return params

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        for gi, group in enumerate(self.param_groups):
            params = self._checked_params(group, True)
            if not params:
                continue
            dev = params[0].device
            for p in params:
                self.state[p]["step"] += 1
            steps = {int(self.state[p]["step"]) for p in (params[0], params[-1])}
            if len(steps) != 1:
                raise RuntimeError("FusedAdamW: parameters of one group must share the step count")
            step_no = steps.pop()
            table, nrec, _ = self._table(gi, params)
            b1, b2 = group["betas"]
            with torch.cuda.device(dev):
                N.check(self._lib.mtgseg_adamw_step(table.data_ptr(), nrec, float(group["lr"]), float(b1), float(b2),
                                                    float(group["eps"]), float(group["weight_decay"]), step_no, None, None,
                                                    N.stream_ptr()), "mtgseg_adamw_step")
            # the update happened outside autograd's view: bump the version counters (what the engine's packed-weight
            # cache keys on); no kernel is launched for this
            torch.autograd.graph.increment_version(params)
        return loss

    # -- captured (CUDA-graph) steps: engine.GraphedTrainStep ---------------------------------------
    def advance(self, hyper):
        """Host half of a captured step, run BEFORE the graph replay: bump the step counts and rewrite the device block
        `hyper` (float32[8]) with this step's lr / betas / eps / weight decay / bias corrections, so that LR schedulers
        (train/train.py:173-181) keep working although the captured launch's scalar arguments are frozen."""
        group = self.param_groups[0]
        steps = [self.state[p]["step"] for p in group["params"] if p in self.state and len(self.state[p])]
        torch._foreach_add_(steps, 1.0)
        b1, b2 = group["betas"]
        with torch.cuda.device(hyper.device):
            N.check(self._lib.mtgseg_adamw_hyper(hyper.data_ptr(), float(group["lr"]), float(b1), float(b2), float(group["eps"]),
                                                 float(group["weight_decay"]), int(steps[0]), N.stream_ptr()), "mtgseg_adamw_hyper")

    @torch.no_grad()
    def step_captured(self, hyper):
        """Device half: the AdamW launch reading its scalars from `hyper`.  Returns what the graph must keep alive."""
        if len(self.param_groups) != 1:
            raise RuntimeError("FusedAdamW: a captured step supports one parameter group (train/train.py:167-171 has one)")
        params = self._checked_params(self.param_groups[0], False)
        entry = self._table(0, params)
        with torch.cuda.device(params[0].device):
            N.check(self._lib.mtgseg_adamw_step_dev(entry[0].data_ptr(), entry[1], hyper.data_ptr(), N.stream_ptr()),
                    "mtgseg_adamw_step_dev")
        return entry
