"""Affine-undo part of ``lightning_pose.data.utils`` (reference ``data/utils.py:142-234``)."""
from __future__ import annotations

import torch

from lightning_pose_b200 import ops

__all__: list[str] = []


def _unit_bbox(n: int, num_views: int, device) -> torch.Tensor:
    # x=0, y=0, h=1, w=1 per view: model_to_frame becomes the identity
    return torch.tensor([0.0, 0.0, 1.0, 1.0], device=device).repeat(n, num_views)


def undo_affine_transform(keypoints: torch.Tensor, transform: torch.Tensor) -> torch.Tensor:
    """(seq, K, 2) keypoints, (2,3) or (seq,2,3) affine -> keypoints with the affine inverted."""
    n, k, _ = keypoints.shape
    flat = keypoints.reshape(n, 2 * k)
    out = ops.remap_keypoints(flat, transform, _unit_bbox(n, 1, flat.device), 1.0, 1.0)
    return out.reshape(n, k, 2)


def undo_affine_transform_batch(
    keypoints_augmented: torch.Tensor, transforms: torch.Tensor, is_multiview: bool = False
) -> torch.Tensor:
    """Undo the augmentation affine when ``transforms.shape[-1] == 3``; otherwise pass through.

    Multiview: ``transforms[v]`` applies to the v-th contiguous block of keypoints (:219-229).
    """
    if transforms.shape[-1] != 3:
        return keypoints_augmented
    n = keypoints_augmented.shape[0]
    v = transforms.shape[0] if is_multiview else 1
    return ops.remap_keypoints(
        keypoints_augmented, transforms, _unit_bbox(n, v, keypoints_augmented.device), 1.0, 1.0,
        is_multiview=is_multiview, num_views=v,
    )
