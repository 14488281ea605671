"""Data-parallel plumbing of the training step (one process per GPU, torch.distributed for the rendezvous).

The hot path shards by image: inference needs no collective at all; training exchanges exactly one thing per step, the
parameter gradients.  The CUDA backward writes all 178 gradients into ONE flat fp32 buffer (``model.last_flat_grad``, 4,201,348
floats = 16.8 MB); the per-parameter ``.grad`` tensors are views of it.  Two ways to average it over the ranks:

* ``enable_gradient_exchange(model)`` (the fast path): the library's own NCCL communicator averages the buffer
  INSIDE ``mtgseg_backward`` in four buckets in reverse execution order, each launched on a communication stream as soon as its
  last gradient exists, overlapped with the rest of the backward pass (csrc/dp_nccl.cu).  ``loss.backward()`` then returns
  averaged gradients, whatever autograd does with the views afterwards, and the step can be captured in a CUDA graph
  (``engine.GraphedTrainStep``).
* ``average_gradients(model)`` after ``loss.backward()``: one ``torch.distributed`` all-reduce of the flat buffer (any backend;
  what the gloo CPU tests exercise).  Not overlapped.

BatchNorm statistics stay per replica (the reference has no SyncBN), i.e. the semantics of DistributedDataParallel around the
reference.
"""
from __future__ import annotations

import torch
import torch.distributed as dist

from . import _native as N


def init_gradient_exchange(group=None) -> int:
    """Collective over the process group: build the library's NCCL communicator on the current CUDA device (the unique id
    travels through ``torch.distributed``).  Returns the world size (1: nothing to do)."""
    if not dist.is_available() or not dist.is_initialized():
        return 1
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    if world == 1:
        return 1
    lib = N.load()
    if lib.mtgseg_dp_world() == world:
        return world
    if lib.mtgseg_dp_world() != 0:
        N.check(lib.mtgseg_dp_shutdown(), "mtgseg_dp_shutdown")
    ident = torch.zeros(128, dtype=torch.uint8)
    if rank == 0:
        N.check(lib.mtgseg_dp_unique_id(ident.data_ptr()), "mtgseg_dp_unique_id")
    on_gpu = dist.get_backend(group) == "nccl"
    t = ident.cuda() if on_gpu else ident
    dist.broadcast(t, src=dist.get_global_rank(group, 0) if group is not None else 0, group=group)
    ident = t.cpu().contiguous()
    N.check(lib.mtgseg_dp_init(ident.data_ptr(), rank, world), "mtgseg_dp_init")
    return world


def enable_gradient_exchange(model, group=None):
    """Make ``loss.backward()`` of ``model`` return rank-averaged gradients (bucketed NCCL all-reduce overlapped with the backward
    pass).  Collective; a no-op for a single process."""
    world = init_gradient_exchange(group)
    model.data_parallel = world > 1
    return model


def average_gradients(target, group=None):
    """In-place mean over the process group of a model's gradients (or of a flat gradient buffer); no-op for a single process.

    For a model, the one all-reduce of ``model.last_flat_grad`` is used only when every ``p.grad`` really is a view of that
    buffer; otherwise (``zero_grad(set_to_none=False)``, gradient accumulation, hooks that clone) the ``.grad`` tensors are
    reduced themselves, so the averaging can never silently miss them."""
    if not dist.is_available() or not dist.is_initialized():
        return target
    world = dist.get_world_size(group)
    if world == 1:
        return target
    if torch.is_tensor(target):
        dist.all_reduce(target, op=dist.ReduceOp.SUM, group=group)
        target.div_(world)
        return target
    model = target
    if getattr(model, "data_parallel", False):
        return model  # already averaged inside backward
    flat = getattr(model, "last_flat_grad", None)
    grads = [p.grad for p in model.parameters() if p.grad is not None]
    if flat is not None and grads and all(g.untyped_storage().data_ptr() == flat.untyped_storage().data_ptr() for g in grads):
        dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
        flat.div_(world)
    else:
        for g in grads:
            dist.all_reduce(g, op=dist.ReduceOp.SUM, group=group)
            g.div_(world)
    return model


def shard_batch(batch_size: int, rank: int, world: int) -> slice:
    """Contiguous shard of a global batch for this rank (remainder goes to the first ranks)."""
    base, rem = divmod(batch_size, world)
    start = rank * base + min(rank, rem)
    return slice(start, start + base + (1 if rank < rem else 0))


def merge_counts(counts: torch.Tensor, group=None) -> torch.Tensor:
    """Global confusion counts of a sharded evaluation: the one (optional) collective of the inference path."""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(counts, op=dist.ReduceOp.SUM, group=group)
    return counts
