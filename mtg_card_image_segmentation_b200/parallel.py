"""Data-parallel plumbing of the training step (one process per GPU, torch.distributed).

The hot path shards by image: inference needs no collective at all; training exchanges exactly one thing per
step, the parameter gradients.  The CUDA backward writes all 178 gradients into ONE flat fp32 buffer
(``model.last_flat_grad``, 4,201,348 floats = 16.8 MB), so the exchange is a single bucket-free all-reduce; the
per-parameter ``.grad`` tensors are views of that buffer and see the averaged values.  BatchNorm statistics stay
per replica (the reference has no SyncBN), i.e. the semantics of DistributedDataParallel around the reference.
"""
from __future__ import annotations

import torch
import torch.distributed as dist


def average_gradients(flat_grad: torch.Tensor, group=None) -> torch.Tensor:
    """In-place mean of the flat gradient buffer over the process group (no-op for a single process)."""
    if not dist.is_available() or not dist.is_initialized():
        return flat_grad
    world = dist.get_world_size(group)
    if world == 1:
        return flat_grad
    dist.all_reduce(flat_grad, op=dist.ReduceOp.SUM, group=group)
    flat_grad.div_(world)
    return flat_grad


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
