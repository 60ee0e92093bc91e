"""Optimizer step of the training hot path: torch.optim.Adam / AdamW semantics in one native launch.

Reference: ``BaseFeatureExtractor.configure_optimizers`` builds ``optim.Adam(params, lr=...)`` or ``optim.AdamW``
(lightning_pose/models/base.py:458-477).  The heatmap head has four parameter tensors (80,971 scalars); torch's fused
multi-tensor kernel covers 65,536 elements per block, so its launch is five blocks and ~38 us long at the very end of
the step.  ``lpb_adam_step`` (csrc/optim.cu) runs a block per 256 elements and keeps the step counter on the device,
so a captured CUDA graph replays it.
"""
from __future__ import annotations

import ctypes as C

import torch

from ._lib import check, lib

__all__ = ["FusedAdam"]

_MAX = 16  # tensors per launch (csrc/optim.cu ADAM_MAX_TENSORS)


class FusedAdam(torch.optim.Optimizer):
    """``torch.optim.Adam`` (``decoupled_weight_decay=False``) or ``AdamW`` (``True``) for CUDA fp32 parameters.

    State per parameter: ``step`` (device float, shared by the parameters of one launch), ``exp_avg``, ``exp_avg_sq``
    -- the keys of torch's own Adam, so checkpoints move between the two.  ``lr`` may be a float or a 0-dim CUDA tensor
    (needed when a learning-rate schedule has to act on a captured graph).  ``amsgrad`` / ``maximize`` are not offered.
    """

    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0, decoupled_weight_decay=False):
        if not 0.0 <= float(betas[0]) < 1.0 or not 0.0 <= float(betas[1]) < 1.0:
            raise ValueError(f"Invalid betas: {betas}")
        if eps < 0.0 or weight_decay < 0.0:
            raise ValueError("eps and weight_decay must be non-negative")
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay, decoupled_weight_decay=decoupled_weight_decay))
        self._arrays = {}

    def _launch(self, group, ps):
        p0 = ps[0]
        for p in ps:
            if not (p.is_cuda and p.dtype == torch.float32 and p.is_contiguous() and p.grad.is_contiguous() and p.grad.dtype == torch.float32):
                raise RuntimeError("FusedAdam: parameters and gradients must be contiguous CUDA fp32 tensors")
            st = self.state[p]
            if not st:
                st["exp_avg"] = torch.zeros_like(p, memory_format=torch.preserve_format)
                st["exp_avg_sq"] = torch.zeros_like(p, memory_format=torch.preserve_format)
        st0 = self.state[p0]
        if "step" not in st0:
            st0["step"] = torch.zeros((), device=p0.device, dtype=torch.float32)
        if "_counter" not in st0:
            st0["_counter"] = torch.zeros((), device=p0.device, dtype=torch.int32)
        step = st0["step"] = st0["step"].to(device=p0.device, dtype=torch.float32)
        for p in ps[1:]:
            self.state[p]["step"] = step  # one counter per launch
        key = tuple(x for p in ps for x in (p.data_ptr(), p.grad.data_ptr(), self.state[p]["exp_avg"].data_ptr(), self.state[p]["exp_avg_sq"].data_ptr()))
        arr = self._arrays.get(key)
        if arr is None:
            n = len(ps)
            vp = C.c_void_p * n
            arr = (vp(*[p.data_ptr() for p in ps]), vp(*[p.grad.data_ptr() for p in ps]), vp(*[self.state[p]["exp_avg"].data_ptr() for p in ps]),
                   vp(*[self.state[p]["exp_avg_sq"].data_ptr() for p in ps]), (C.c_int64 * n)(*[p.numel() for p in ps]))
            self._arrays = {key: arr}  # the pointers of a training run do not change; keep the latest set only
        lr = group["lr"]
        lr_dev = lr.data_ptr() if isinstance(lr, torch.Tensor) else None
        b1, b2 = group["betas"]
        with torch.cuda.device(p0.device):
            check(lib.lpb_adam_step(len(ps), arr[0], arr[1], arr[2], arr[3], arr[4], step.data_ptr(), st0["_counter"].data_ptr(),
                                    0.0 if lr_dev else float(lr), lr_dev, float(b1), float(b2), float(group["eps"]), float(group["weight_decay"]),
                                    int(bool(group["decoupled_weight_decay"])), torch.cuda.current_stream(p0.device).cuda_stream))

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        for group in self.param_groups:
            ps = [p for p in group["params"] if p.grad is not None]
            for i in range(0, len(ps), _MAX):
                self._launch(group, ps[i:i + _MAX])
        return loss
