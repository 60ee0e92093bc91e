"""Video-ingest boundary: decoded frames -> ``UnlabeledBatchDict``.

Mirror of what ``LitDaliWrapper._dali_output_to_tensors`` hands the trackers
(``lightning_pose/data/video/dali.py:265-327``): ``frames`` (seq, 3, H, W) normalised with the ImageNet statistics,
``transforms`` (a lone ``[-1]`` when no geometric augmentation ran, :175-178), ``bbox`` = ``[0, 0, height, width]`` of the
ORIGINAL frame repeated per frame, ``is_multiview``.  Video decoding itself (NVDEC / DALI readers, :133-156) is host/driver
I/O and out of scope: the entry point takes decoded uint8 RGB surfaces already on the device and fuses resize + /255 +
normalise + layout change into one kernel (``csrc/ingest.cu``).
"""
from __future__ import annotations

from typing import Sequence

import torch

from lightning_pose_b200 import ops

__all__ = ["frames_to_unlabeled_batch", "context_windows_step"]


def frames_to_unlabeled_batch(frames_u8: torch.Tensor | Sequence[torch.Tensor], resize_dims: Sequence[int] | None = None,
                              dtype: torch.dtype = torch.float32, channels_last: bool = False) -> dict:
    """uint8 (seq, H, W, 3) frames of one view - or a list of them, one per view - to the batch dict of the trackers.

    Single view -> ``UnlabeledBatchDict``; several views -> ``MultiviewUnlabeledBatchDict`` with frames
    (seq, views, 3, H, W), transforms (views, 1), bbox (seq, 4 * views) (reference :289-327).
    """
    views = [frames_u8] if isinstance(frames_u8, torch.Tensor) else list(frames_u8)
    outs, boxes = [], []
    for v in views:
        outs.append(ops.frames_normalize(v, size=resize_dims, channels_last=channels_last, dtype=dtype))
        boxes.append(torch.tensor([0.0, 0.0, float(v.shape[1]), float(v.shape[2])], device=v.device))
    seq = outs[0].shape[0]
    if len(views) == 1:
        return {"frames": outs[0], "transforms": torch.tensor([-1.0], device=outs[0].device), "bbox": boxes[0].repeat(seq, 1), "is_multiview": False}
    return {
        "frames": torch.stack(outs, dim=1),
        "transforms": torch.full((len(views), 1), -1.0, device=outs[0].device),
        "bbox": torch.cat(boxes).repeat(seq, 1),
        "is_multiview": True,
    }


def context_windows_step(sequence_length: int) -> int:
    """Reader step of context (MHCRNN) prediction: consecutive sequences overlap by 4 frames (dali.py:214-216)."""
    return int(sequence_length) - 4
