"""Data-parallel plumbing for the hot path: one process per GPU, exactly one collective per step.

The reference scales with Lightning DDP (``lightning_pose/train.py:411-428``): per-rank labeled batch =
ceil(B/N), per-rank clip length = ceil(T/N), context = ceil((T-4)/N)+4 (``data/factory.py:250-285``).
Its three implicit collective sources (DDP gradient buckets, SyncBN, one ``sync_dist`` all-reduce per
logged scalar, ``models/base.py:535-544``) collapse here into flat-buffer ``all_reduce`` buckets: ``[head gradients |
loss scalars]`` - the logged scalars ride the tail of the gradient buffer - plus, when a backbone trains along, its own
bucket launched early on a side stream so it overlaps the backward.
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
    """Bucketed flat-buffer all-reduce(mean) of parameter gradients plus a few logged scalars.

    Two ways to drive it:

    * ``begin_step()`` ... backward ... ``finish_step(scalars)`` - the training loop's form.  ``begin_step`` zeroes the
      head bucket and points every ``p.grad`` at its slice, so autograd accumulates straight into the flat buffer
      (no pack / unpack copies); if a backbone bucket exists (``extra_floats``) its all-reduce is launched right away on
      a side stream and overlaps the backward.  ``finish_step`` appends the scalars, joins the side stream and issues
      the head bucket's all-reduce; ``p.grad`` then holds the averaged gradients in place.
    * ``step(scalars)`` - one-shot form for gradients that already live in their own ``p.grad`` tensors (packs,
      reduces, scatters back; a rank whose ``p.grad`` is None contributes zeros AND receives the average, so replicas
      cannot drift apart when a loss branch is inactive on some ranks).

    Works on any backend (NCCL on GPUs with ``ReduceOp.AVG``; gloo in the CPU tests with SUM + divide).
    """

    def __init__(self, params: Iterable[torch.nn.Parameter], n_scalars: int = 0, group=None, extra_floats: int = 0) -> None:
        self.params = [p for p in params if p.requires_grad]
        self.n_scalars = int(n_scalars)
        self.group = group
        self.sizes = [p.numel() for p in self.params]
        self.total = sum(self.sizes) + self.n_scalars
        ref = self.params[0]
        self.buffer = torch.zeros(self.total, dtype=torch.float32, device=ref.device)
        self.views = []
        off = 0
        for p, n in zip(self.params, self.sizes):
            self.views.append(self.buffer[off : off + n].view_as(p))
            off += n
        self.extra = torch.zeros(int(extra_floats), dtype=torch.float32, device=ref.device) if extra_floats else None
        self.side = torch.cuda.Stream(device=ref.device) if (ref.is_cuda and self.extra is not None) else None
        self.launches = 0

    # ---- helpers ----------------------------------------------------------------------------------
    def _world(self) -> int:
        return dist.get_world_size(self.group) if dist.is_initialized() else 1

    def _mean_all_reduce(self, t: torch.Tensor) -> None:
        world = self._world()
        if world <= 1:
            return
        if t.is_cuda:
            dist.all_reduce(t, op=dist.ReduceOp.AVG, group=self.group)
        else:
            dist.all_reduce(t, op=dist.ReduceOp.SUM, group=self.group)
            t.div_(world)
        self.launches += 1

    # ---- training-loop form -------------------------------------------------------------------------
    def begin_step(self) -> None:
        self.buffer.zero_()
        for p, v in zip(self.params, self.views):
            p.grad = v
        if self.extra is not None:
            if self.side is not None:
                self.side.wait_stream(torch.cuda.current_stream(self.extra.device))
                with torch.cuda.stream(self.side):
                    self._mean_all_reduce(self.extra)
            else:
                self._mean_all_reduce(self.extra)

    def finish_step(self, scalars: Sequence[torch.Tensor] = ()) -> torch.Tensor:
        if len(scalars) != self.n_scalars:
            raise ValueError(f"expected {self.n_scalars} scalars, got {len(scalars)}")
        tail = self.buffer[self.total - self.n_scalars :]
        if self.n_scalars:
            tail.copy_(torch.stack([s.detach().reshape(()).float() for s in scalars]))
        self._mean_all_reduce(self.buffer)
        if self.side is not None:
            torch.cuda.current_stream(self.extra.device).wait_stream(self.side)
        return tail

    # ---- one-shot form ------------------------------------------------------------------------------
    def step(self, scalars: Sequence[torch.Tensor] = ()) -> torch.Tensor:
        if len(scalars) != self.n_scalars:
            raise ValueError(f"expected {self.n_scalars} scalars, got {len(scalars)}")
        for p, v in zip(self.params, self.views):
            if p.grad is None:
                v.zero_()
            elif p.grad.data_ptr() != v.data_ptr():
                v.copy_(p.grad)
        tail = self.buffer[self.total - self.n_scalars :]
        if self.n_scalars:
            tail.copy_(torch.stack([s.detach().reshape(()).float() for s in scalars]))
        self._mean_all_reduce(self.buffer)
        for p, v in zip(self.params, self.views):
            if p.grad is None:
                p.grad = v.clone()  # this rank had no gradient: it still receives the average
            elif p.grad.data_ptr() != v.data_ptr():
                p.grad.copy_(v)
        return tail.clone()
