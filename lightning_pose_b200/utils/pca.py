"""KeypointPCA: the PCA-loss parameter holder, mirror of ``lightning_pose.utils.pca.KeypointPCA``.

On the hot path only the device-side arithmetic matters - ``_format_data`` (reference
``lightning_pose/utils/pca.py:97-190``), ``reproject`` (:266-294) and
``compute_reprojection_error`` (:296-309) - and ``PCALoss.__call__`` evaluates all of it inside one
CUDA kernel (``csrc/losses.cu``).  The one-off CPU *fit* (NaN-aware covariance + eigh, component
selection, empirical epsilon; reference :311-328, :331-470, :611-756) is out of scope for the
accelerated path; a compact numpy version is provided so a loss can be built from a
keypoint array, but reference-fitted parameters can equally be assigned to ``.parameters``.
"""
from __future__ import annotations

import warnings
from typing import Literal

import numpy as np
import torch

from lightning_pose_b200 import ops

__all__ = ["KeypointPCA", "format_multiview_data_for_pca"]


def format_multiview_data_for_pca(data_arr: torch.Tensor, mirrored_column_matches) -> torch.Tensor:
    """(batch, K, 2) -> (batch * K_mv, 2 * n_views): one row per (frame, body part), columns
    ``[x_view0, y_view0, x_view1, ...]`` (reference :759-792)."""
    blocks = []
    n_per_view = len(mirrored_column_matches[0])
    for cols in mirrored_column_matches:
        assert len(cols) == n_per_view
        idx = torch.as_tensor(np.asarray(cols), dtype=torch.long, device=data_arr.device)
        blocks.append(data_arr.index_select(1, idx).reshape(-1, 2))
    return torch.cat(blocks, dim=1)


class KeypointPCA:
    """Collects labeled keypoints, fits PCA once on the host, serves parameters to the loss kernel."""

    def __init__(
        self,
        loss_type: Literal["pca_singleview", "pca_multiview"],
        data_module=None,
        components_to_keep: int | float | None = 0.99,
        empirical_epsilon_percentile: float = 99.0,
        mirrored_column_matches=None,
        columns_for_singleview_pca=None,
        device: str | torch.device = "cpu",
        centering_method: Literal["mean", "median"] | None = None,
    ) -> None:
        self.loss_type = loss_type
        self.data_module = data_module
        self.components_to_keep = components_to_keep
        self.empirical_epsilon_percentile = empirical_epsilon_percentile
        if mirrored_column_matches is not None and isinstance(mirrored_column_matches[0], int):
            # flat per-view index list: expand over the dataset's views (reference :76-88)
            dataset = getattr(data_module, "dataset", None)
            view_names = getattr(dataset, "view_names", None)
            if not view_names:
                raise ValueError(
                    "cfg.data.mirrored_column_matches must contain a list of indices for each mirrored view"
                )
            per_view = int(dataset.num_keypoints) // len(view_names)
            mirrored_column_matches = [
                (v * per_view + np.asarray(mirrored_column_matches, dtype=int)).tolist() for v in range(len(view_names))
            ]
        self.mirrored_column_matches = mirrored_column_matches
        self.columns_for_singleview_pca = columns_for_singleview_pca
        self.pca_object = None
        self.device = device
        self.centering_method = centering_method
        self.parameters: dict[str, torch.Tensor] = {}

    # ---- formatting (device-side, data movement only) -------------------------------------------
    def _multiview_format(self, data_arr: torch.Tensor) -> torch.Tensor:
        assert self.mirrored_column_matches is not None
        return format_multiview_data_for_pca(
            data_arr.reshape(data_arr.shape[0], -1, 2), self.mirrored_column_matches
        )

    def _singleview_format(self, data_arr: torch.Tensor) -> torch.Tensor:
        kp = data_arr.reshape(data_arr.shape[0], -1, 2)
        if self.columns_for_singleview_pca is not None:
            idx = torch.as_tensor(np.asarray(self.columns_for_singleview_pca), dtype=torch.long, device=kp.device)
            kp = kp.index_select(1, idx)
        if self.centering_method == "mean":
            kp = kp - kp.mean(dim=1, keepdim=True)
        elif self.centering_method == "median":
            kp = kp - kp.quantile(dim=1, q=0.5, keepdim=True)
        elif self.centering_method is not None:
            raise NotImplementedError(f"centering_method: {self.centering_method}")
        return kp.reshape(kp.shape[0], -1)

    def _format_data(self, data_arr: torch.Tensor) -> torch.Tensor:
        if self.loss_type == "pca_multiview":
            return self._multiview_format(data_arr)
        if self.loss_type == "pca_singleview":
            return self._singleview_format(data_arr)
        raise NotImplementedError(self.loss_type)

    # ---- what the fused kernel consumes ----------------------------------------------------------
    def kernel_params(self, num_keypoints: int, epsilon: float | torch.Tensor) -> ops.PcaParams:
        """Pack (selection indices, mean, kept eigenvectors, epsilon) for ``ops.unsup_losses``."""
        mean = self.parameters["mean"]
        kept = self.parameters["kept_eigenvectors"]
        if self.loss_type == "pca_multiview":
            mcm = self.mirrored_column_matches
            idx = np.asarray(mcm, dtype=np.int32).reshape(-1)
            return ops.PcaParams(idx, len(mcm[0]), len(mcm), None, mean, kept, float(epsilon), mean.device)
        cols = self.columns_for_singleview_pca
        idx = np.arange(num_keypoints, dtype=np.int32) if cols is None else np.asarray(cols, dtype=np.int32)
        return ops.PcaParams(idx, len(idx), 0, self.centering_method, mean, kept, float(epsilon), mean.device)

    def reproject(self, data_arr: torch.Tensor | None = None) -> torch.Tensor:
        """((x - mu) V^T) V + mu for already-formatted rows (reference :266-294).  Off the fused path
        (diagnostics / metrics only): two small device matmuls."""
        if data_arr is None:
            data_arr = self.data_arr.to(self.device)
        evecs = self.parameters["kept_eigenvectors"]
        mean = self.parameters["mean"].unsqueeze(0)
        assert data_arr.shape[1] == evecs.shape[1] == mean.shape[1] and data_arr.shape[1] % 2 == 0
        return ((data_arr - mean) @ evecs.T) @ evecs + mean

    def compute_reprojection_error(self, data_arr: torch.Tensor | None = None) -> torch.Tensor:
        """Per-keypoint reprojection error (rows, D/2) of formatted rows (reference :296-309)."""
        if data_arr is None:
            data_arr = self.data_arr.to(self.device)
        diff = (data_arr - self.reproject(data_arr)).reshape(data_arr.shape[0], -1, 2)
        return torch.linalg.norm(diff, dim=2)

    # ---- host-side fit (one-off init; see module docstring) --------------------------------------
    def _get_data(self) -> None:
        dm = self.data_module
        if isinstance(dm, (torch.Tensor, np.ndarray)):
            arr = dm
        elif hasattr(dm, "labeled_keypoints"):
            arr = dm.labeled_keypoints
        else:
            raise TypeError(
                "KeypointPCA fit needs the labeled training keypoints: pass a (N, 2K) tensor/array or an object "
                "with a `labeled_keypoints` attribute (the reference's DataExtractor is data-layer, out of scope)"
            )
        self.data_arr = torch.as_tensor(np.asarray(arr), dtype=torch.float32)

    def _check_data(self) -> None:
        if self.data_arr.shape[0] < self.data_arr.shape[1]:
            raise ValueError(
                f"cannot fit PCA with {self.data_arr.shape[0]} samples < {self.data_arr.shape[1]} observation dimensions"
            )

    def _fit_pca(self) -> None:
        x = self.data_arr.numpy().astype(np.float64)
        mean = np.nanmean(x, axis=0)
        cov = np.ma.cov(np.ma.masked_invalid(x), rowvar=False).data  # NaN-aware covariance
        evals, evecs = np.linalg.eigh(cov)
        order = np.argsort(evals)[::-1]
        evals, evecs = np.clip(evals[order], 0.0, None), evecs[:, order].T
        # deterministic sign: largest-|.| entry of each component positive (sklearn svd_flip, v-based)
        signs = np.sign(evecs[np.arange(evecs.shape[0]), np.argmax(np.abs(evecs), axis=1)])
        evecs = evecs * signs[:, None]
        self.pca_object = {"mean_": mean, "components_": evecs, "explained_variance_ratio_": evals / evals.sum()}

    def _choose_n_components(self) -> None:
        n_all = self.pca_object["components_"].shape[0]
        if self.loss_type == "pca_multiview":
            self._n_components_kept = 3  # (x, y, z) explains every view
            if self.components_to_keep != 3:
                warnings.warn(
                    f"for {self.loss_type} loss, you specified {self.components_to_keep} components_to_keep, "
                    "but we will instead keep 3 components",
                    stacklevel=2,
                )
            return
        keep = self.components_to_keep
        if keep is None:
            n = n_all
        elif type(keep) is int:
            if keep > n_all:
                raise ValueError(f"components_to_keep was set to {keep}, exceeding the maximum value of {n_all} observation dims")
            n = keep
        else:
            if not 0.0 <= keep <= 1.0:
                raise ValueError(f"components_to_keep was set to {keep} while it has to be between 0.0 and 1.0")
            cum = np.cumsum(self.pca_object["explained_variance_ratio_"])
            n = n_all if keep >= cum[-1] else int(np.where(cum >= keep)[0][0]) + 1
        self._n_components_kept = int(n)

    def _set_parameter_dict(self) -> None:
        k = self._n_components_kept
        comps = self.pca_object["components_"]
        self.parameters = {
            "mean": torch.tensor(self.pca_object["mean_"], dtype=torch.float, device=self.device),
            "kept_eigenvectors": torch.tensor(comps[:k], dtype=torch.float, device=self.device),
            "discarded_eigenvectors": torch.tensor(comps[k:], dtype=torch.float, device=self.device),
        }
        err = self.compute_reprojection_error().detach().cpu().numpy().reshape(-1)
        eps = float(np.nanpercentile(err, self.empirical_epsilon_percentile, axis=0))
        self.parameters["epsilon"] = torch.tensor(eps, dtype=torch.float, device=self.device)

    def __call__(self) -> None:
        self._get_data()
        self.data_arr = self._format_data(self.data_arr)
        self._check_data()
        self._fit_pca()
        self._choose_n_components()
        self._set_parameter_dict()
