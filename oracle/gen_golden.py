"""Generate tests/golden/*.npz by executing the reference's own files.  TEST INFRASTRUCTURE.

Run in the authoring container only:  ``python -m oracle.gen_golden``
(needs /root/reference; the GPU box never runs this - it consumes the committed fixtures).

Every array stored under ``out_*`` was produced by the *reference's* code
(``lightning_pose/models/heads/heatmap.py``, ``data/heatmaps.py``, ``losses/losses.py``,
``losses/factory.py``, ``utils/pca.py``, ``data/utils.py``, ``data/bboxes.py``) loaded unmodified
through ``oracle/ref_loader.py``; ``in_*`` arrays are the seeded inputs.  The script also asserts
that ``oracle/lp_oracle.py`` reproduces each output (this is what "oracle pinned" means).
"""
from __future__ import annotations

import math
import os

import numpy as np
import torch

from oracle import lp_oracle as O
from oracle import ref_loader as R

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def _np(x):
    return x.detach().cpu().numpy() if isinstance(x, torch.Tensor) else np.asarray(x)


def _check(name, ref, mine, atol=1e-6, rtol=1e-5):
    ref, mine = _np(ref), _np(mine)
    assert ref.shape == mine.shape, (name, ref.shape, mine.shape)
    ok = np.allclose(ref, mine, atol=atol, rtol=rtol, equal_nan=True)
    err = float(np.nanmax(np.abs(ref - mine))) if ref.size else 0.0
    print(f"  oracle-vs-reference {name:40s} max|diff| = {err:.3e} {'OK' if ok else 'MISMATCH'}")
    assert ok, name


def gen_decode(hm, dh):
    g = {}
    # --- reference known-answer tests: tests/models/heads/test_heatmap.py:126-219 -------------
    for ds in (1, 2, 3):
        x = torch.zeros(1, 4, 8, 8)
        x[0, 0, 2, 2] = 1.0
        x[0, 1, 4, 4] = 1.0
        x[0, 2, 0, 0] = 1.0
        x[0, 3, 0, 1] = 1.0
        p, c = hm.run_subpixelmaxima(x, ds, torch.tensor(1000.0))
        g[f"kat_ds{ds}_in"] = _np(x)
        g[f"kat_ds{ds}_out_preds"] = _np(p)
        g[f"kat_ds{ds}_out_conf"] = _np(c)
        po, co = O.decode_softargmax(x, ds, 1000.0)
        _check(f"decode KAT ds={ds}", p, po)
        _check(f"decode KAT conf ds={ds}", c, co)
    x = torch.zeros(1, 1, 8, 8)
    x[0, 0, 4, 4] = 1.0
    for temp in (1000.0, 100.0, 10.0):
        p, c = hm.run_subpixelmaxima(x, 2, torch.tensor(temp))
        g[f"temp{int(temp)}_out_preds"] = _np(p)
        g[f"temp{int(temp)}_out_conf"] = _np(c)
        po, co = O.decode_softargmax(x, 2, temp)
        _check(f"decode T={temp}", p, po)
        _check(f"decode conf T={temp}", c, co)
    g["temp_in"] = _np(x)

    # --- SURVEY A.3 golden vectors at the real size (96x96 -> 384x384) -------------------------
    kp = torch.tensor([[[100.3, 200.7], [5.2, 380.1], [float("nan"), float("nan")], [383.9, 0.1], [192.0, 192.0]]])
    t = dh.generate_heatmaps(kp, 384, 384, (96, 96))
    p, c = hm.run_subpixelmaxima(t, 2, torch.tensor(1000.0))
    g["a3_in_keypoints"] = _np(kp)
    g["a3_out_targets_sum"] = _np(t.sum((2, 3)))
    g["a3_out_targets_peak"] = _np(t.amax((2, 3)))
    g["a3_out_preds"] = _np(p)
    g["a3_out_conf"] = _np(c)
    po, co = O.decode_softargmax(O.gaussian_targets(kp, 384, 384, (96, 96)), 2, 1000.0)
    _check("A.3 decode", p, po, atol=2e-5)
    _check("A.3 conf", c, co)

    # --- seeded regimes: peaked / flat / edge / multimodal / negative-valued, non-square ------
    gen = torch.Generator().manual_seed(4)
    kp = torch.rand(3, 6, 2, generator=gen) * torch.tensor([128.0 - 16, 96.0 - 16]) + 8
    peaked = dh.generate_heatmaps(kp, 96, 128, (24, 32))
    peaked = peaked + 1e-6
    peaked = peaked / peaked.sum((2, 3), keepdim=True)
    flat = torch.softmax(0.01 * torch.randn(2, 6, 24 * 32, generator=torch.Generator().manual_seed(5)), -1).reshape(2, 6, 24, 32)
    ekp = torch.tensor([[[0.5, 0.5], [127.0, 95.0], [1.0, 60.0], [126.0, 3.0], [64.0, 0.2], [0.0, 95.9]]])
    edge = dh.generate_heatmaps(ekp, 96, 128, (24, 32))
    multi = 0.55 * peaked[:1] + 0.45 * torch.flip(peaked[:1], dims=(2, 3))
    raw = torch.randn(1, 6, 24, 32, generator=torch.Generator().manual_seed(8)) * 0.02  # final_softmax=False style
    raw[0, :, 10, 20] += 0.3
    cases = {"peaked": peaked, "flat": flat, "edge": edge, "multi": multi, "raw": raw}
    for name, h in cases.items():
        for ds in (1, 2, 3):
            p, c = hm.run_subpixelmaxima(h.clone(), ds, torch.tensor(1000.0))
            g[f"{name}_ds{ds}_out_preds"] = _np(p)
            g[f"{name}_ds{ds}_out_conf"] = _np(c)
            po, co = O.decode_softargmax(h, ds, 1000.0)
            _check(f"decode {name} ds={ds}", p, po, atol=2e-5)
            _check(f"decode conf {name} ds={ds}", c, co)
        g[f"{name}_in"] = _np(h)

    # --- upsample operator: impulse responses of the reference's `upsample` --------------------
    for n, ds in ((16, 1), (16, 2), (12, 3), (96, 2)):
        # plane c = e_c 1^T  ->  field = U[:, c] (x) s with s = U 1;  ones plane -> s (x) s
        probe = torch.eye(n).reshape(1, n, n, 1).repeat(1, 1, 1, n)
        ones_plane = torch.ones(1, 1, n, n)
        for _ in range(ds):
            probe = hm.upsample(probe)
            ones_plane = hm.upsample(ones_plane)
        jmid = (n << ds) // 2
        s_mid = float(ones_plane[0, 0, jmid, jmid]) ** 0.5
        u_ref = _np(probe[0, :, :, jmid]).T / s_mid
        g[f"U_n{n}_ds{ds}"] = u_ref.astype(np.float32)
        u_mine = O.upsample_matrix_1d(n, ds)
        _check(f"upsample matrix n={n} ds={ds}", u_ref, u_mine, atol=2e-6)
        s = u_mine.sum(1)
        _check(f"ones response n={n} ds={ds}", _np(ones_plane[0, 0]), np.outer(s, s), atol=2e-6)
    return g


def gen_targets(dh):
    g = {}
    # visibility x NaN x out-of-bounds matrix (tests/data/test_heatmaps.py:203-454)
    kp = torch.tensor(
        [[[10.0, 20.0], [float("nan"), float("nan")], [-30.0, 5.0], [40.0, 300.0], [63.9, 47.9], [-2.0, -2.0]],
         [[0.0, 0.0], [31.5, 23.25], [70.0, 10.0], [10.0, float("nan")], [64.0, 48.0], [66.5, 50.6]]]
    )
    g["in_keypoints"] = _np(kp)
    t = dh.generate_heatmaps(kp, 48, 64, (12, 16))
    g["out_vis_none"] = _np(t)
    _check("targets vis=None", t, O.gaussian_targets(kp, 48, 64, (12, 16)))
    vis = torch.tensor([[2, 2, 2, 1, 0, 2], [1, 0, 2, 2, 2, 2]])
    g["in_visibility"] = _np(vis)
    t = dh.generate_heatmaps(kp, 48, 64, (12, 16), visibility=vis)
    g["out_vis"] = _np(t)
    _check("targets vis", t, O.gaussian_targets(kp, 48, 64, (12, 16), visibility=vis))
    t = dh.generate_heatmaps(kp, 48, 64, (24, 32), sigma=2.0)
    g["out_sigma2_ds1"] = _np(t)
    _check("targets sigma=2", t, O.gaussian_targets(kp, 48, 64, (24, 32), sigma=2.0))
    # evaluate_heatmaps_at_location (tests/data/test_heatmaps.py:457-563)
    gen = torch.Generator().manual_seed(11)
    hmaps = torch.rand(2, 3, 10, 14, generator=gen)
    locs = torch.tensor([[[0.0, 0.0], [13.9, 9.9], [6.5, 4.5]], [[1.2, 8.7], [12.0, 0.3], [7.0, 7.0]]])
    v = dh.evaluate_heatmaps_at_location(hmaps, locs)
    g["eval_in_heatmaps"] = _np(hmaps)
    g["eval_in_locs"] = _np(locs)
    g["eval_out"] = _np(v)
    _check("evaluate_heatmaps_at_location", v, O.confidence_window_sum(hmaps, locs))
    return g


def gen_head(hm):
    g = {}
    torch.manual_seed(21)
    # resnet-style: stride 32, ds=2 -> 2 deconv layers; ViT-style: stride 16, ds=2 -> 1 layer
    for tag, arch, cin, k, hh, ww in (("resnet", "resnet50", 64, 5, 3, 4), ("vit", "vits_dino", 32, 5, 6, 4)):
        head = hm.HeatmapHead(arch, cin, k)
        for layer in list(head.upsampling_layers)[1:]:
            torch.nn.init.xavier_uniform_(layer.weight, gain=3.0)  # peaked logits
            torch.nn.init.uniform_(layer.bias, -0.5, 0.5)
        feats = torch.randn(2, cin, hh, ww) * 0.5
        out = head(feats)
        g[f"{tag}_in_features"] = _np(feats)
        ws, bs = [], []
        for i, layer in enumerate(list(head.upsampling_layers)[1:]):
            g[f"{tag}_w{i}"] = _np(layer.weight)
            g[f"{tag}_b{i}"] = _np(layer.bias)
            ws.append(layer.weight.detach())
            bs.append(layer.bias.detach())
        g[f"{tag}_out_heatmaps"] = _np(out)
        _check(f"head {tag}", out, O.head_forward(feats, ws, bs))
        head.final_softmax = False
        out = head(feats)
        g[f"{tag}_out_logits"] = _np(out)
        _check(f"head logits {tag}", out, O.head_forward(feats, ws, bs, final_softmax=False), atol=1e-5)
    return g


def gen_remap(du, db):
    g = {}
    gen = torch.Generator().manual_seed(31)
    t_frames, k = 6, 4
    kp = torch.rand(t_frames, 2 * k, generator=gen) * 100
    ang = math.radians(7.0)
    a = torch.tensor([[1.1 * math.cos(ang), -1.1 * math.sin(ang), 3.0], [1.1 * math.sin(ang), 1.1 * math.cos(ang), -2.0]])
    g["in_keypoints"] = _np(kp)
    g["in_transform_shared"] = _np(a)
    out = du.undo_affine_transform_batch(kp.clone(), a, False)
    g["out_affine_shared"] = _np(out)
    _check("undo affine shared", out, O.undo_affine(kp, a))
    per = a[None].repeat(t_frames, 1, 1) + 0.05 * torch.randn(t_frames, 2, 3, generator=gen)
    g["in_transform_perframe"] = _np(per)
    out = du.undo_affine_transform_batch(kp.clone(), per, False)
    g["out_affine_perframe"] = _np(out)
    _check("undo affine per frame", out, O.undo_affine(kp, per))
    mv = torch.stack([a, a * 0.9 + 0.1])
    g["in_transform_multiview"] = _np(mv)
    out = du.undo_affine_transform_batch(kp.clone(), mv, True)
    g["out_affine_multiview"] = _np(out)
    _check("undo affine multiview", out, O.undo_affine(kp, mv, True))
    out = du.undo_affine_transform_batch(kp.clone(), torch.ones(1), False)
    _check("undo affine identity", out, O.undo_affine(kp, torch.ones(1)))
    # model_to_frame_batch: unlabeled single view, context (bbox has 4 extra rows), multiview
    bbox = torch.tensor([[5.0, 7.0, 396.0, 406.0]]).repeat(t_frames, 1) + torch.arange(t_frames)[:, None]
    frames = torch.zeros(t_frames, 3, 128, 256)
    out = db.model_to_frame_batch({"frames": frames, "bbox": bbox, "is_multiview": False}, kp.clone())
    g["in_bbox"] = _np(bbox)
    g["out_frame_single"] = _np(out)
    _check("model_to_frame single", out, O.model_to_frame(kp, bbox, 128, 256))
    bbox_ctx = torch.cat([bbox[:2] * 0 + 1, bbox, bbox[:2] * 0 + 2])
    out = db.model_to_frame_batch({"frames": frames, "bbox": bbox_ctx, "is_multiview": False}, kp.clone())
    g["in_bbox_ctx"] = _np(bbox_ctx)
    g["out_frame_ctx"] = _np(out)
    _check("model_to_frame context", out, O.model_to_frame(kp, bbox_ctx, 128, 256))
    bbox_mv = torch.cat([bbox, bbox * 0.5 + 3], dim=1)
    out = db.model_to_frame_batch({"frames": frames, "bbox": bbox_mv, "is_multiview": True}, kp.clone())
    g["in_bbox_mv"] = _np(bbox_mv)
    g["out_frame_mv"] = _np(out)
    _check("model_to_frame multiview", out, O.model_to_frame(kp, bbox_mv, 128, 256, num_views=2))
    return g


def gen_losses(ll, lf, dh, pca_mod):
    g = {}
    # --- TemporalLoss: SURVEY A.3 / tests/losses/test_losses.py:343-392 -----------------------
    kp = torch.tensor([[10, 10, 50, 50], [13, 14, 50, 50], [13, 14, 90, 50], [40, 14, 90, 80], [40, 14, 90, 80], [41, 14, 90, 80]], dtype=torch.float32)
    conf = torch.tensor([[0.9, 0.9], [0.9, 0.01], [0.9, 0.9], [0.9, 0.9], [0.01, 0.9], [0.9, 0.9]])
    tl = ll.TemporalLoss(epsilon=[2.0, 20.0], prob_threshold=0.05)
    v1, _ = tl(kp, conf)
    v2, _ = tl(kp)
    g["temporal_in_kp"], g["temporal_in_conf"] = _np(kp), _np(conf)
    g["temporal_out_conf"], g["temporal_out_noconf"] = _np(v1), _np(v2)
    _check("temporal conf", v1, O.temporal_loss(kp, conf, [2.0, 20.0], 0.05))
    _check("temporal noconf", v2, O.temporal_loss(kp, None, [2.0, 20.0], 0.05))
    gen = torch.Generator().manual_seed(41)
    kp2 = torch.cumsum(torch.randn(32, 34, generator=gen) * 8, 0) + 200
    conf2 = torch.rand(32, 17, generator=gen)
    tl2 = ll.TemporalLoss(epsilon=20.0, prob_threshold=0.05)
    v3, _ = tl2(kp2, conf2)
    g["temporal2_in_kp"], g["temporal2_in_conf"], g["temporal2_out"] = _np(kp2), _np(conf2), _np(v3)
    _check("temporal T=32", v3, O.temporal_loss(kp2, conf2, 20.0, 0.05))

    # --- heatmap losses: SURVEY A.3 ------------------------------------------------------------
    a = dh.generate_heatmaps(torch.tensor([[[100.0, 100.0], [200.0, 300.0]]]), 384, 384, (96, 96))
    b = dh.generate_heatmaps(torch.tensor([[[104.0, 100.0], [float("nan"), 1.0]]]), 384, 384, (96, 96))
    g["hm_in_a_kp"] = np.array([[[100.0, 100.0], [200.0, 300.0]]], np.float32)
    g["hm_in_b_kp"] = np.array([[[104.0, 100.0], [np.nan, 1.0]]], np.float32)
    for nm, cls, fn in (("mse", ll.HeatmapMSELoss, O.heatmap_mse_loss), ("kl", ll.HeatmapKLLoss, O.heatmap_kl_loss), ("js", ll.HeatmapJSLoss, O.heatmap_js_loss)):
        v_ba, _ = cls()(heatmaps_targ=b, heatmaps_pred=a)
        v_ab, _ = cls()(heatmaps_targ=a, heatmaps_pred=b)
        g[f"hm_{nm}_out_targb_preda"], g[f"hm_{nm}_out_targa_predb"] = _np(v_ba), _np(v_ab)
        _check(f"heatmap {nm} (targ=b)", v_ba, fn(b, a))
        _check(f"heatmap {nm} (targ=a)", v_ab, fn(a, b), rtol=1e-4)
    # seeded batch with dropped planes, softmax-like predictions
    gen = torch.Generator().manual_seed(42)
    kpl = torch.rand(4, 6, 2, generator=gen) * 128
    kpl[0, 1] = float("nan")
    kpl[2, 4] = float("nan")
    vis = torch.tensor([[2, 2, 1, 2, 0, 2]]).repeat(4, 1)
    targ = dh.generate_heatmaps(kpl, 128, 128, (32, 32), visibility=vis)
    pred = torch.softmax(torch.randn(4, 6, 32 * 32, generator=gen) * 2.0, -1).reshape(4, 6, 32, 32)
    g["hmb_in_kp"], g["hmb_in_vis"], g["hmb_in_pred"] = _np(kpl), _np(vis), _np(pred)
    for nm, cls, fn in (("mse", ll.HeatmapMSELoss, O.heatmap_mse_loss), ("kl", ll.HeatmapKLLoss, O.heatmap_kl_loss), ("js", ll.HeatmapJSLoss, O.heatmap_js_loss)):
        v, _ = cls()(heatmaps_targ=targ, heatmaps_pred=pred)
        g[f"hmb_{nm}_out"] = _np(v)
        _check(f"heatmap batch {nm}", v, fn(targ, pred))

    # --- TemporalHeatmapLoss ---------------------------------------------------------------
    hseq = torch.softmax(torch.randn(6, 3, 16 * 16, generator=gen) * 3.0, -1).reshape(6, 3, 16, 16)
    cseq = torch.rand(6, 3, generator=gen)
    g["thm_in_heatmaps"], g["thm_in_conf"] = _np(hseq), _np(cseq)
    for kind in ("mse", "kl"):
        eps = 1e-5 if kind == "mse" else [0.5, 1.0, 2.0]
        th = ll.TemporalHeatmapLoss(loss_name=f"temporal_heatmap_{kind}", epsilon=eps, prob_threshold=0.2)
        v, _ = th(hseq.clone(), cseq)
        g[f"thm_{kind}_out"] = _np(v)
        _check(f"temporal heatmap {kind}", v, O.temporal_heatmap_loss(hseq, cseq, kind, eps, 0.2))

    # --- PCA losses with synthetic parameters (fit is out of scope; params are inputs) -------
    def make_pca_loss(loss_name, params, eps, **kw):
        pca = pca_mod.KeypointPCA(loss_type=loss_name, data_module=object(), device="cpu", **kw)
        pca.parameters = params
        loss = object.__new__(ll.PCALoss)
        ll.Loss.__init__(loss, log_weight=5.0)
        loss.device, loss.loss_name, loss.pca = "cpu", loss_name, pca
        loss.epsilon = torch.tensor(eps, dtype=torch.float)
        return loss

    gen = torch.Generator().manual_seed(43)
    kseq = torch.cumsum(torch.randn(32, 34, generator=gen) * 3, 0) + 150
    cols = [0, 1, 2, 3, 5, 6, 8, 9, 10, 12, 13, 14, 15, 16]
    d = 2 * len(cols)
    q, _ = torch.linalg.qr(torch.randn(d, d, generator=gen))
    kept = q[:6].contiguous()
    mean = torch.rand(d, generator=gen) * 300
    g["pca_in_kp"], g["pca_sv_cols"] = _np(kseq), np.array(cols)
    g["pca_sv_mean"], g["pca_sv_kept"] = _np(mean), _np(kept)
    for centering in (None, "mean", "median"):
        loss = make_pca_loss("pca_singleview", {"mean": mean, "kept_eigenvectors": kept}, 2.5,
                             columns_for_singleview_pca=cols, centering_method=centering)
        v, _ = loss(kseq)
        g[f"pca_sv_out_{centering}"] = _np(v)
        fm = O.pca_format_singleview(kseq, cols, centering)
        _check(f"pca singleview centering={centering}", v, O.pca_loss(fm, mean, kept, 2.5))
    mcm = [[0, 1, 2, 3, 4, 5, 6], [8, 9, 10, 11, 12, 13, 14]]
    q4, _ = torch.linalg.qr(torch.randn(4, 4, generator=gen))
    kept4, mean4 = q4[:3].contiguous(), torch.rand(4, generator=gen) * 300
    loss = make_pca_loss("pca_multiview", {"mean": mean4, "kept_eigenvectors": kept4}, 0.7, mirrored_column_matches=mcm)
    v, _ = loss(kseq)
    g["pca_mv_mcm"], g["pca_mv_mean"], g["pca_mv_kept"], g["pca_mv_out"] = np.array(mcm), _np(mean4), _np(kept4), _np(v)
    _check("pca multiview", v, O.pca_loss(O.pca_format_multiview(kseq, mcm), mean4, kept4, 0.7))

    # --- ReprojectionHeatmapLoss (the slot a re-created UnimodalLoss would occupy) -----------
    rl = ll.ReprojectionHeatmapLoss(128, 128, 32, 32, log_weight=1.0)
    kr = kpl + torch.randn(4, 6, 2, generator=gen) * 3
    v, _ = rl(heatmaps_targ=targ, keypoints_pred_2d_reprojected=kr)
    g["reproj_in_kp"], g["reproj_out"] = _np(kr), _np(v)
    _check("reprojection heatmap", v, O.reprojection_heatmap_loss(targ, kr, 128, 128, (32, 32)))

    # --- LossFactory total: weights, anneal exemption (losses/factory.py:229-285) -----------
    fac = lf.LossFactory({"heatmap_mse": {"log_weight": 0.0}, "temporal": {"log_weight": 5.0, "epsilon": 20.0, "prob_threshold": 0.05}}, None)
    tot, logs = fac(stage="train", anneal_weight=0.3, heatmaps_targ=targ, heatmaps_pred=pred, keypoints_pred=kp2, confidences=conf2)
    g["factory_out_total"] = _np(tot)
    g["factory_log_names"] = np.array([d_["name"] for d_ in logs])
    g["factory_log_values"] = np.array([float(d_["value"]) for d_ in logs], np.float32)
    mine = O.combine_losses({"heatmap_mse": (O.heatmap_mse_loss(targ, pred), 0.0), "temporal": (O.temporal_loss(kp2, conf2, 20.0, 0.05), 5.0)}, 0.3)
    _check("loss factory total", tot, mine)
    return g


def _mhcrnn_params(head):
    m = head.head_mf
    p = {"W_f": (m.W_f.weight, m.W_f.bias), "W_b": (m.W_b.weight, m.W_b.bias),
         "H_f": (m.H_f[0].weight, m.H_f[0].bias, m.H_f[1].weight, m.H_f[1].bias),
         "H_b": (m.H_b[0].weight, m.H_b[0].bias, m.H_b[1].weight, m.H_b[1].bias)}
    if m.upsampling_factor == 2:
        p["W_pre"] = (m.W_pre.weight, m.W_pre.bias)
    return p


def gen_mhcrnn(mh, base_src_path):
    """HeatmapMHCRNNHead / UpsamplingCRNN (heads/heatmap_mhcrnn.py) and get_context_from_sequence (models/base.py:159-196;
    base.py itself needs Lightning to import, so that one function is executed from its source text)."""
    import ast
    import textwrap

    g = {}
    src = open(base_src_path).read()
    fn = next(n for n in ast.parse(src).body if isinstance(n, ast.FunctionDef) and n.name == "get_context_from_sequence")
    code = "\n".join(src.splitlines()[fn.lineno - 1 : fn.end_lineno])
    code = textwrap.dedent(code)
    # strip the jaxtyping annotations of the signature (not importable pieces of base.py's header)
    body_start = code.index('"""')
    code = "def get_context_from_sequence(img_seq, context_length):\n    " + code[body_start:]
    ns = {"torch": torch}
    exec(code, ns)
    gen = torch.Generator().manual_seed(51)
    seq = torch.randn(7, 3, 4, 4, generator=gen)
    out = ns["get_context_from_sequence"](seq, 5)
    g["ctx_in_seq"], g["ctx_out_windows"] = _np(seq), _np(out)
    _check("get_context_from_sequence", out, O.context_windows(seq, 5))
    for tag, arch, cin, uf in (("vit", "vits_dino", 64, 1), ("resnet", "resnet50", 64, 2)):
        torch.manual_seed(52)
        head = mh.HeatmapMHCRNNHead(arch, cin, 5, upsampling_factor=uf)
        for prm in head.head_sf.parameters():  # peaked, non-degenerate maps
            torch.nn.init.normal_(prm, std=0.3)
        feats = torch.randn(3, cin, 4, 6, 5, generator=gen)
        sf, mf = head(feats, torch.Size([3, 5, 3, 64, 96]), False)
        g[f"{tag}_in_features"] = _np(feats)
        for name, t in head.state_dict().items():
            if ".layers." not in name:  # ModuleList aliases of the same tensors
                g[f"{tag}_param_{name}"] = _np(t)
        g[f"{tag}_out_sf"], g[f"{tag}_out_mf"] = _np(sf), _np(mf)
        _check(f"mhcrnn multi-frame {tag}", mf, O.mhcrnn_multiframe(feats.permute(4, 0, 1, 2, 3), _mhcrnn_params(head), uf))
        d = list(head.head_sf.upsampling_layers)[1:]
        _check(f"mhcrnn single-frame {tag}", sf, O.head_forward(feats[..., 2], [x.weight for x in d], [x.bias for x in d]))
    return g


def main():
    assert R.reference_available(), "needs /root/reference (authoring container only)"
    os.makedirs(GOLDEN_DIR, exist_ok=True)
    hm = R.load("lightning_pose.models.heads.heatmap")
    dh = R.load("lightning_pose.data.heatmaps")
    ll = R.load("lightning_pose.losses.losses")
    lf = R.load("lightning_pose.losses.factory")
    du = R.load("lightning_pose.data.utils")
    db = R.load("lightning_pose.data.bboxes")
    pm = R.load("lightning_pose.utils.pca")
    mh = R.load("lightning_pose.models.heads.heatmap_mhcrnn")
    with torch.no_grad():
        groups = {
            "decode": gen_decode(hm, dh),
            "targets": gen_targets(dh),
            "head": gen_head(hm),
            "remap": gen_remap(du, db),
            "losses": gen_losses(ll, lf, dh, pm),
            "mhcrnn": gen_mhcrnn(mh, os.path.join(R.REF_ROOT, "lightning_pose", "models", "base.py")),
        }
    for name, arrays in groups.items():
        path = os.path.join(GOLDEN_DIR, f"{name}.npz")
        np.savez_compressed(path, **arrays)
        print(f"wrote {path}: {len(arrays)} arrays, {os.path.getsize(path) / 1024:.1f} KiB")


if __name__ == "__main__":
    main()
