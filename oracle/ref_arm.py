"""The CPU arm of ``bench.py``: BASELINE configs[1]'s hot path on the REFERENCE'S OWN code.  TEST INFRASTRUCTURE.

Every numerical step below is a call into the reference's unmodified files (``models/heads/heatmap.py``,
``data/heatmaps.py``, ``data/utils.py``, ``data/bboxes.py``, ``losses/losses.py``, ``utils/pca.py``), executed through
``oracle/ref_loader.py`` from ``/root/reference`` (authoring container) or the staged ``oracle/_ref`` (GPU box; recipe:
``oracle/build_ref.py``).  Only ``kornia`` (absent from the image) is restated (ref_loader).  Nothing here imports
``lightning_pose_b200``.  Falls back to the restated oracle (``kind = "port"``) only if no reference copy is reachable.
"""
from __future__ import annotations

import numpy as np
import torch

from oracle import lp_oracle as O
from oracle import ref_loader as R


class ReferenceStep:
    """forward (+ autograd backward) of one semi-supervised step, as ``HeatmapTracker`` composes it
    (``models/heatmap_tracker.py:107-191, :264-340``): labeled and unlabeled batches go through the head separately."""

    def __init__(self, prob: dict, img: int, hm: int, b_labeled: int, t_unlabeled: int):
        self.kind = "reference" if R.reference_available() else "port"
        self.prob, self.img, self.hm, self.bl, self.tu = prob, img, hm, b_labeled, t_unlabeled
        w1, b1, w2, b2 = (t.detach().clone().float() for t in prob["head_params"])
        self.params = [w1, b1, w2, b2]
        if self.kind == "reference":
            hmod = R.load("lightning_pose.models.heads.heatmap")
            self.dh = R.load("lightning_pose.data.heatmaps")
            self.du = R.load("lightning_pose.data.utils")
            self.db = R.load("lightning_pose.data.bboxes")
            ll = R.load("lightning_pose.losses.losses")
            pm = R.load("lightning_pose.utils.pca")
            k = w2.shape[1]
            self.head = hmod.HeatmapHead("resnet50", 4 * w1.shape[0], k)
            d1, d2 = list(self.head.upsampling_layers)[1:]
            with torch.no_grad():
                d1.weight.copy_(w1), d1.bias.copy_(b1), d2.weight.copy_(w2), d2.bias.copy_(b2)
            self.params = [d1.weight, d1.bias, d2.weight, d2.bias]
            self.mse = ll.HeatmapMSELoss()
            self.temporal = ll.TemporalLoss(epsilon=20.0, prob_threshold=0.05, log_weight=5.0)
            pca = pm.KeypointPCA(loss_type="pca_singleview", data_module=object(), device="cpu")
            pca.parameters = {"mean": prob["pca"]["mean"], "kept_eigenvectors": prob["pca"]["kept"]}
            self.pca = object.__new__(ll.PCALoss)
            ll.Loss.__init__(self.pca, log_weight=5.0)
            self.pca.device, self.pca.loss_name, self.pca.pca = "cpu", "pca_singleview", pca
            self.pca.epsilon = torch.tensor(prob["pca"]["eps"], dtype=torch.float)

    def __call__(self, train: bool = True) -> torch.Tensor:
        p, n = self.prob, self.prob["n_clips"]
        nl = n * self.bl
        w_unsup = 1.0 / (2.0 * np.exp(5.0))
        feats = p["feats"].detach().float()
        f_lab, f_unl = feats[:nl].requires_grad_(train), feats[nl:].requires_grad_(train)
        for t in self.params:
            t.grad = None
            t.requires_grad_(train)
        with torch.set_grad_enabled(train):
            if self.kind == "reference":
                hm_lab = self.head(f_lab)
                _kp_lab, _cf_lab = self.head.run_subpixelmaxima(hm_lab)
                targ = self.dh.generate_heatmaps(p["kp_lab"], self.img, self.img, (self.hm, self.hm), visibility=p["vis"])
                l_sup, _ = self.mse(heatmaps_targ=targ, heatmaps_pred=hm_lab)
                hm_unl = self.head(f_unl)
                kp, cf = self.head.run_subpixelmaxima(hm_unl)
                kp = self.du.undo_affine_transform_batch(kp, p["tf"], False)
                kp = self.db.model_to_frame_batch({"frames": torch.zeros(1, 1, self.img, self.img).expand(n * self.tu, 3, self.img, self.img), "bbox": p["bbox"], "is_multiview": False}, kp)
                tot = 0.5 * l_sup
                for c in range(n):  # each clip is its own unlabeled batch (base.py:627-658)
                    sl = slice(c * self.tu, (c + 1) * self.tu)
                    lt, _ = self.temporal(kp[sl], cf[sl].detach())
                    lp, _ = self.pca(kp[sl])
                    tot = tot + (lt + lp) * w_unsup
            else:
                w1, b1, w2, b2 = self.params
                hm_lab = O.head_forward(f_lab, [w1, w2], [b1, b2])
                O.decode_softargmax(hm_lab.detach(), 2, 1000.0)
                targ = O.gaussian_targets(p["kp_lab"], self.img, self.img, (self.hm, self.hm), visibility=p["vis"])
                l_sup = O.heatmap_mse_loss(targ, hm_lab)
                kp, cf = O.decode_softargmax(O.head_forward(f_unl, [w1, w2], [b1, b2]), 2, 1000.0)
                kp = O.model_to_frame(O.undo_affine(kp, p["tf"]), p["bbox"], self.img, self.img)
                tot = 0.5 * l_sup
                for c in range(n):
                    sl = slice(c * self.tu, (c + 1) * self.tu)
                    tot = tot + (O.temporal_loss(kp[sl], cf[sl].detach(), 20.0, 0.05)
                                 + O.pca_loss(O.pca_format_singleview(kp[sl]), p["pca"]["mean"], p["pca"]["kept"], p["pca"]["eps"])) * w_unsup
        if train:
            tot.backward()
        return tot.detach()
