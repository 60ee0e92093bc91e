"""CPU: pin oracle/lp_oracle.py against the golden vectors produced by the reference's own code
(oracle/gen_golden.py) and against the reference's known-answer tests."""
import numpy as np
import pytest
import torch

from oracle import lp_oracle as O

T = torch.from_numpy


def close(a, b, atol=1e-6, rtol=1e-5):
    a = a.detach().numpy() if isinstance(a, torch.Tensor) else np.asarray(a)
    np.testing.assert_allclose(a, np.asarray(b), atol=atol, rtol=rtol, equal_nan=True)


# ---- reference known answers: tests/models/heads/test_heatmap.py:126-219 ----------------------
@pytest.mark.parametrize("ds", [1, 2])
def test_kat_interior_points(ds):
    x = torch.zeros(1, 2, 8, 8)
    x[0, 0, 2, 2] = 1.0
    x[0, 1, 4, 4] = 1.0
    preds, conf = O.decode_softargmax(x, ds, 1000.0)
    close(preds[0], [2.0 * 2**ds, 2.0 * 2**ds, 4.0 * 2**ds, 4.0 * 2**ds], atol=1e-5)
    assert conf[0, 0] == 1.0 and conf[0, 1] == 1.0


@pytest.mark.parametrize("ds", [1, 2])
def test_kat_boundary_points(ds):
    x = torch.zeros(1, 2, 8, 8)
    x[0, 0, 0, 0] = 1.0
    x[0, 1, 0, 1] = 1.0
    preds, conf = O.decode_softargmax(x, ds, 1000.0)
    assert abs(preds[0, 0]) <= 0.5 and abs(preds[0, 1]) <= 0.5
    assert abs(preds[0, 2] - 2**ds) <= 0.1 * 2**ds and abs(preds[0, 3]) <= 0.5
    close(conf[0], [1.0, 1.0], rtol=1e-3)


def test_kat_temperature():
    x = torch.zeros(1, 1, 8, 8)
    x[0, 0, 4, 4] = 1.0
    p, c = O.decode_softargmax(x, 2, 1000.0)
    close(p[0], [16.0, 16.0], atol=1e-5)
    assert c[0, 0] == 1.0
    p, c = O.decode_softargmax(x, 2, 100.0)
    close(p[0], [16.0, 16.0], atol=1e-4)
    assert c[0, 0] != 1.0 and abs(c[0, 0] - 1.0) < 1e-3
    p, c = O.decode_softargmax(x, 2, 10.0)
    assert c[0, 0] < 0.5


# ---- SURVEY Appendix A.3 (hand-checkable table) --------------------------------------------------
def test_a3_table():
    kp = torch.tensor([[[100.3, 200.7], [5.2, 380.1], [float("nan")] * 2, [383.9, 0.1], [192.0, 192.0]]])
    t = O.gaussian_targets(kp, 384, 384, (96, 96))
    close(t.sum((2, 3))[0], [1, 1, 0, 1, 1], atol=2e-6)
    close(t.amax((2, 3))[0], [0.1006844, 0.1630782, 0.0, 0.3236310, 0.1018592], atol=1e-6)
    p, c = O.decode_softargmax(t, 2, 1000.0)
    close(p[0], [100.3026505, 200.6925201, 5.3018045, 378.4868774, 190.0, 189.9999847,
                 379.4949341, 1.5076725, 192.0000153, 192.0], atol=2e-4)
    close(c[0], [0.9990522, 0.9999651, 25 / 147456, 1.0, 0.9996977], atol=2e-6)


# ---- golden vectors from the reference's own code ----------------------------------------------
def test_decode_golden(golden):
    g = golden("decode")
    for ds in (1, 2, 3):
        p, c = O.decode_softargmax(T(g[f"kat_ds{ds}_in"]), ds, 1000.0)
        close(p, g[f"kat_ds{ds}_out_preds"], atol=1e-5)
        close(c, g[f"kat_ds{ds}_out_conf"])
    for temp in (1000, 100, 10):
        p, c = O.decode_softargmax(T(g["temp_in"]), 2, float(temp))
        close(p, g[f"temp{temp}_out_preds"], atol=1e-5)
        close(c, g[f"temp{temp}_out_conf"])
    for name in ("peaked", "flat", "edge", "multi", "raw"):
        for ds in (1, 2, 3):
            p, c = O.decode_softargmax(T(g[f"{name}_in"]), ds, 1000.0)
            close(p, g[f"{name}_ds{ds}_out_preds"], atol=2e-5)
            close(c, g[f"{name}_ds{ds}_out_conf"])


@pytest.mark.parametrize("n,ds", [(16, 1), (16, 2), (12, 3), (96, 2)])
def test_upsample_matrix_golden(golden, n, ds):
    u = O.upsample_matrix_1d(n, ds)
    close(u, golden("decode")[f"U_n{n}_ds{ds}"], atol=2e-6)
    if (n, ds) == (96, 2):  # SURVEY A.1 properties
        assert (np.abs(u) > 0).sum(1).max() <= 9
        close(u[160, 36:44], [8.48e-5, -6.2872e-3, -1.222e-4, 0.4202645, 0.5465790, 0.0525428, -0.0131401, 7.83e-5], atol=1e-6)
        close(u.sum(1)[:6], [0.48805, 0.72321, 0.86133, 0.94327, 0.98642, 1.00031], atol=1e-5)
        assert abs(np.abs(u).sum(1).max() - 1.08496) < 1e-5


def test_targets_golden(golden):
    g = golden("targets")
    kp, vis = T(g["in_keypoints"]), T(g["in_visibility"])
    close(O.gaussian_targets(kp, 48, 64, (12, 16)), g["out_vis_none"])
    close(O.gaussian_targets(kp, 48, 64, (12, 16), visibility=vis), g["out_vis"])
    close(O.gaussian_targets(kp, 48, 64, (24, 32), sigma=2.0), g["out_sigma2_ds1"])
    close(O.confidence_window_sum(T(g["eval_in_heatmaps"]), T(g["eval_in_locs"])), g["eval_out"], atol=2e-6)


@pytest.mark.parametrize("tag,nl", [("resnet", 2), ("vit", 1)])
def test_head_golden(golden, tag, nl):
    g = golden("head")
    ws = [T(g[f"{tag}_w{i}"]) for i in range(nl)]
    bs = [T(g[f"{tag}_b{i}"]) for i in range(nl)]
    f = T(g[f"{tag}_in_features"])
    close(O.head_forward(f, ws, bs), g[f"{tag}_out_heatmaps"], atol=1e-7)
    close(O.head_forward(f, ws, bs, final_softmax=False), g[f"{tag}_out_logits"], atol=1e-5)


def test_remap_golden(golden):
    g = golden("remap")
    kp = T(g["in_keypoints"])
    close(O.undo_affine(kp, T(g["in_transform_shared"])), g["out_affine_shared"], atol=1e-5)
    close(O.undo_affine(kp, T(g["in_transform_perframe"])), g["out_affine_perframe"], atol=1e-5)
    close(O.undo_affine(kp, T(g["in_transform_multiview"]), True), g["out_affine_multiview"], atol=1e-5)
    close(O.undo_affine(kp, torch.ones(1)), g["in_keypoints"])
    close(O.model_to_frame(kp, T(g["in_bbox"]), 128, 256), g["out_frame_single"], atol=1e-5)
    close(O.model_to_frame(kp, T(g["in_bbox_ctx"]), 128, 256), g["out_frame_ctx"], atol=1e-5)
    close(O.model_to_frame(kp, T(g["in_bbox_mv"]), 128, 256, num_views=2), g["out_frame_mv"], atol=1e-5)


def test_losses_golden(golden):
    g = golden("losses")
    kp, conf = T(g["temporal_in_kp"]), T(g["temporal_in_conf"])
    assert abs(float(O.temporal_loss(kp, conf, [2.0, 20.0], 0.05)) - 3.8) < 1e-6
    assert abs(float(O.temporal_loss(kp, None, [2.0, 20.0], 0.05)) - 5.8) < 1e-6
    close(O.temporal_loss(kp, conf, [2.0, 20.0], 0.05), g["temporal_out_conf"])
    close(O.temporal_loss(T(g["temporal2_in_kp"]), T(g["temporal2_in_conf"]), 20.0, 0.05), g["temporal2_out"])
    a = O.gaussian_targets(T(g["hm_in_a_kp"]), 384, 384, (96, 96))
    b = O.gaussian_targets(T(g["hm_in_b_kp"]), 384, 384, (96, 96))
    for nm, fn in (("mse", O.heatmap_mse_loss), ("kl", O.heatmap_kl_loss), ("js", O.heatmap_js_loss)):
        close(fn(b, a), g[f"hm_{nm}_out_targb_preda"])
        close(fn(a, b), g[f"hm_{nm}_out_targa_predb"], rtol=1e-4)
    close(O.heatmap_mse_loss(b, a), 0.0150606, atol=1e-6)
    close(O.heatmap_kl_loss(a, b), 10.0308437, rtol=1e-5)
    targ = O.gaussian_targets(T(g["hmb_in_kp"]), 128, 128, (32, 32), visibility=T(g["hmb_in_vis"]))
    pred = T(g["hmb_in_pred"])
    for nm, fn in (("mse", O.heatmap_mse_loss), ("kl", O.heatmap_kl_loss), ("js", O.heatmap_js_loss)):
        close(fn(targ, pred), g[f"hmb_{nm}_out"])
    hseq, cseq = T(g["thm_in_heatmaps"]), T(g["thm_in_conf"])
    close(O.temporal_heatmap_loss(hseq, cseq, "mse", 1e-5, 0.2), g["thm_mse_out"])
    close(O.temporal_heatmap_loss(hseq, cseq, "kl", [0.5, 1.0, 2.0], 0.2), g["thm_kl_out"])
    kseq = T(g["pca_in_kp"])
    cols = g["pca_sv_cols"].tolist()
    for centering in (None, "mean", "median"):
        fm = O.pca_format_singleview(kseq, cols, centering)
        close(O.pca_loss(fm, T(g["pca_sv_mean"]), T(g["pca_sv_kept"]), 2.5), g[f"pca_sv_out_{centering}"])
    fm = O.pca_format_multiview(kseq, g["pca_mv_mcm"].tolist())
    close(O.pca_loss(fm, T(g["pca_mv_mean"]), T(g["pca_mv_kept"]), 0.7), g["pca_mv_out"])
    close(O.reprojection_heatmap_loss(targ, T(g["reproj_in_kp"]), 128, 128, (32, 32)), g["reproj_out"])
    tot = O.combine_losses(
        {"heatmap_mse": (O.heatmap_mse_loss(targ, pred), 0.0),
         "temporal": (O.temporal_loss(T(g["temporal2_in_kp"]), T(g["temporal2_in_conf"]), 20.0, 0.05), 5.0)}, 0.3)
    close(tot, g["factory_out_total"])


# ---- more of the reference's own numerical unit tests, restated as known answers for the oracle ---------------
def test_kat_confidence_window_exact_sums():
    """tests/data/test_heatmaps.py:457-563: the window sum is the exact sum of the (2r+1)^2 patch around
    (trunc y, trunc x), r = floor(1.25 * 2) = 2, zero beyond the plane's border."""
    p = torch.zeros(1, 2, 12, 12)
    p[0, 0, 5, 6] = 0.25
    p[0, 0, 7, 8] = 0.5      # inside the 5x5 window centred on (y=5, x=6): rows 3..7, cols 4..8
    p[0, 0, 8, 6] = 0.125    # row 8: outside
    p[0, 1, 0, 0] = 0.75     # corner: window reaches beyond the plane, missing part counts as zero
    p[0, 1, 2, 2] = 0.0625
    p[0, 1, 3, 0] = 1.0      # row 3: outside the window centred on (0, 0)
    locs = torch.tensor([[[6.9, 5.2], [0.4, 0.9]]])  # (x, y); trunc -> (6, 5) and (0, 0)
    close(O.confidence_window_sum(p, locs), [[0.75, 0.8125]])


def test_kat_heatmap_losses_zero_at_equality_and_monotone():
    """tests/losses/test_losses.py:137-217, :449-480: MSE / KL / JS vanish for identical heatmaps and grow as the
    prediction is rolled away from the target."""
    g = torch.Generator().manual_seed(0)
    targ = O.gaussian_targets(torch.rand(3, 4, 2, generator=g) * 40 + 12, 64, 64, (32, 32))
    for fn in (O.heatmap_mse_loss, O.heatmap_kl_loss, O.heatmap_js_loss):
        assert abs(float(fn(targ, targ.clone()))) < 1e-6
        vals = [float(fn(targ, torch.roll(targ, s, dims=-1))) for s in (1, 2, 4)]
        assert vals[0] > 1e-6 and vals[0] < vals[1] < vals[2]


def test_kat_temporal_per_keypoint_epsilon():
    """tests/losses/test_losses.py:343-392: analytic values with a scalar and with a per-keypoint epsilon."""
    kp = torch.zeros(3, 4)          # T=3 frames, K=2 keypoints
    kp[1] = torch.tensor([3.0, 4.0, 0.0, 0.0])   # keypoint 0 moves by 5 then by 5 back; keypoint 1 rests
    close(O.temporal_loss(kp), 2.5)                                   # mean(5, 0, 5, 0)
    close(O.temporal_loss(kp, epsilon=1.0), 2.0)                      # mean(4, 0, 4, 0)
    close(O.temporal_loss(kp, epsilon=[4.0, 0.0]), 0.5)               # mean(1, 0, 1, 0)
    conf = torch.tensor([[0.9, 0.9], [0.01, 0.9], [0.9, 0.9]])        # frame 1 of keypoint 0 is low confidence
    close(O.temporal_loss(kp, conf, epsilon=0.0, prob_threshold=0.05), 0.0)  # both of its differences are masked


def test_kat_pca_zero_inside_subspace_and_positive_outside():
    """tests/losses/test_losses.py:289-311: points inside the kept subspace reproject onto themselves (loss 0);
    moving along a discarded direction gives the distance, rectified by epsilon."""
    g = torch.Generator().manual_seed(1)
    q, _ = torch.linalg.qr(torch.randn(6, 6, generator=g))
    kept, disc = q[:2], q[2:]
    mean = torch.randn(6, generator=g)
    inside = mean[None] + torch.randn(5, 2, generator=g) @ kept
    close(O.pca_loss(inside, mean, kept, 0.0), 0.0, atol=1e-6)
    off = inside + 2.0 * disc[0][None]
    err = O.pca_reprojection_error(off, mean, kept)                    # (N, 3): norm of the residual per (x, y) pair
    close((err**2).sum(1), torch.full((5,), 4.0), atol=1e-5)           # the residual is exactly 2 * disc[0]
    assert float(O.pca_loss(off, mean, kept, 0.0)) > float(O.pca_loss(off, mean, kept, 0.5)) > 0.0
    close(O.pca_loss(off, mean, kept, 10.0), 0.0)


def test_kat_factory_anneal_exemption():
    """tests/losses/test_factory.py:213-263: the anneal weight multiplies every loss except heatmap_{mse,kl,js}."""
    one = torch.tensor(1.0)
    losses = {"heatmap_mse": (one, 0.0), "temporal": (one, 0.0), "pca_singleview": (2 * one, np.log(2.0))}
    w = O.loss_weight(0.0)  # 1 / (2 * exp(0)) = 0.5
    close(O.combine_losses(losses, 1.0), w + w + 2 * O.loss_weight(np.log(2.0)))
    close(O.combine_losses(losses, 0.0), w)           # only the heatmap loss survives anneal_weight = 0
    close(O.combine_losses(losses, 0.25), w + 0.25 * (w + 2 * O.loss_weight(np.log(2.0))))


def test_mhcrnn_oracle_matches_reference_goldens(golden):
    """heads/heatmap_mhcrnn.py (UpsamplingCRNN / HeatmapMHCRNNHead) and models/base.py:159-196 (context windows)."""
    g = golden("mhcrnn")
    close(O.context_windows(T(g["ctx_in_seq"]), 5), g["ctx_out_windows"])
    for tag, uf in (("vit", 1), ("resnet", 2)):
        pr = lambda n: T(g[f"{tag}_param_head_mf.{n}"])
        p = {"W_f": (pr("W_f.weight"), pr("W_f.bias")), "W_b": (pr("W_b.weight"), pr("W_b.bias")),
             "H_f": (pr("H_f.0.weight"), pr("H_f.0.bias"), pr("H_f.1.weight"), pr("H_f.1.bias")),
             "H_b": (pr("H_b.0.weight"), pr("H_b.0.bias"), pr("H_b.1.weight"), pr("H_b.1.bias"))}
        if uf == 2:
            p["W_pre"] = (pr("W_pre.weight"), pr("W_pre.bias"))
        feats = T(g[f"{tag}_in_features"])
        close(O.mhcrnn_multiframe(feats.permute(4, 0, 1, 2, 3), p, uf), g[f"{tag}_out_mf"], atol=1e-8)
        n = 1 if uf == 1 else 2
        ws = [T(g[f"{tag}_param_head_sf.upsampling_layers.{i + 1}.weight"]) for i in range(n)]
        bs = [T(g[f"{tag}_param_head_sf.upsampling_layers.{i + 1}.bias"]) for i in range(n)]
        close(O.head_forward(feats[..., 2], ws, bs), g[f"{tag}_out_sf"], atol=1e-8)
