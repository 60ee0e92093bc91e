"""Data-parallel plumbing for the hot path: one process per GPU, exactly one collective per step.

The reference scales with Lightning DDP (``lightning_pose/train.py:411-428``): per-rank labeled batch =
ceil(B/N), per-rank clip length = ceil(T/N), context = ceil((T-4)/N)+4 (``data/factory.py:250-285``).
Its three implicit collective sources (DDP gradient buckets, SyncBN, one ``sync_dist`` all-reduce per
logged scalar, ``models/base.py:535-544``) collapse here into ONE ``all_reduce`` over a flat buffer:
``[head gradients | loss scalars]`` - the logged scalars ride the tail of the gradient buffer.
Clips are independent units, so the forward/backward data path itself needs no communication.
"""
from __future__ import annotations

import math
from typing import Iterable, Sequence

import torch
import torch.distributed as dist

__all__ = ["per_rank_sizes", "FlatGradAllReducer"]


def per_rank_sizes(train_batch_size: int, sequence_length: int, context_batch_size: int, world_size: int) -> dict:
    """Per-GPU sizes that keep the effective batch constant (reference ``data/factory.py:250-285``)."""
    n = max(int(world_size), 1)
    return {
        "train_batch_size": int(math.ceil(train_batch_size / n)),
        "sequence_length": int(math.ceil(sequence_length / n)),
        "context_batch_size": int(math.ceil(max(context_batch_size - 4, 0) / n + 4)),
    }


class FlatGradAllReducer:
    """Single flat-buffer all-reduce(mean) of parameter gradients plus a few scalars.

    ``step(scalars)`` packs every ``p.grad`` and the given 0-dim tensors into one contiguous buffer,
    issues one ``dist.all_reduce``, divides by the world size and scatters the results back
    (gradients in place; returns the averaged scalars).  Works on any backend (NCCL on GPUs, gloo in
    the CPU tests).
    """

    def __init__(self, params: Iterable[torch.nn.Parameter], n_scalars: int = 0, group=None) -> None:
        self.params = [p for p in params if p.requires_grad]
        self.n_scalars = int(n_scalars)
        self.group = group
        self.sizes = [p.numel() for p in self.params]
        self.total = sum(self.sizes) + self.n_scalars
        ref = self.params[0]
        self.buffer = torch.zeros(self.total, dtype=torch.float32, device=ref.device)
        self.launches = 0

    def step(self, scalars: Sequence[torch.Tensor] = ()) -> torch.Tensor:
        if len(scalars) != self.n_scalars:
            raise ValueError(f"expected {self.n_scalars} scalars, got {len(scalars)}")
        off = 0
        for p, n in zip(self.params, self.sizes):
            if p.grad is None:
                self.buffer[off : off + n].zero_()
            else:
                self.buffer[off : off + n].copy_(p.grad.reshape(-1))
            off += n
        for s in scalars:
            self.buffer[off].copy_(s.detach().reshape(()))
            off += 1
        world = dist.get_world_size(self.group) if dist.is_initialized() else 1
        if world > 1:
            dist.all_reduce(self.buffer, op=dist.ReduceOp.SUM, group=self.group)
            self.launches += 1
            self.buffer.div_(world)
        off = 0
        for p, n in zip(self.params, self.sizes):
            if p.grad is not None:
                p.grad.copy_(self.buffer[off : off + n].view_as(p.grad))
            off += n
        return self.buffer[off : off + self.n_scalars].clone()
