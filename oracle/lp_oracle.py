"""CPU oracle for the lightning-pose hot path.  TEST INFRASTRUCTURE ONLY.

This module is a plain torch-CPU fp32 restatement of the reference algorithm for the one
hot path this repository accelerates (heatmap head -> Gaussian targets -> soft-argmax decode
-> coordinate remap -> loss stack).  It is the *checker*:

  * only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
    ``--impl reference`` leg may import it;
  * nothing under ``lightning_pose_b200/`` imports it and the product path has no CPU fallback.

Pinning ("parity pinned"): ``oracle/gen_golden.py`` executes the reference's *own* files
(``/root/reference/lightning_pose/{models/heads/heatmap.py,data/heatmaps.py,losses/losses.py,
losses/factory.py,utils/pca.py,data/utils.py,data/bboxes.py}``) through a stub loader and writes
the golden vectors in ``tests/golden/``; ``tests/test_oracle_golden.py`` checks every function
below against those vectors and against the reference's known-answer tests
(``tests/models/heads/test_heatmap.py:126-219``, ``tests/losses/test_losses.py:343-392`` ...).

Third-party arithmetic: the reference calls ``kornia`` (unpinned, ``pyproject.toml:40``; absent
from this image).  The six kornia functions on the path are restated here from their published
definitions and are pinned by the reference's known-answer tests above.

Every function cites the reference ``file:line`` it follows (paths relative to /root/reference).
"""
from __future__ import annotations

import math

import numpy as np
import torch
import torch.nn.functional as F

# --------------------------------------------------------------------------------------
# kornia restatements (kornia.geometry.subpix / kornia.filters / kornia.losses)
# --------------------------------------------------------------------------------------


def spatial_softmax2d(x: torch.Tensor, temperature: float | torch.Tensor = 1.0) -> torch.Tensor:
    """softmax over the flattened (H, W) plane of ``temperature * x``.

    kornia.geometry.subpix.spatial_softmax2d; call sites lightning_pose/models/heads/heatmap.py:126,211.
    """
    b, c, h, w = x.shape
    t = torch.as_tensor(temperature, dtype=x.dtype)
    flat = x.reshape(b, c, h * w) * t
    return torch.softmax(flat, dim=-1).reshape(b, c, h, w)


def spatial_expectation2d(p: torch.Tensor) -> torch.Tensor:
    """(sum p*col, sum p*row) in pixel units -> (B, C, 2) ordered (x, y).

    kornia.geometry.subpix.spatial_expectation2d(normalized_coordinates=False);
    call site lightning_pose/models/heads/heatmap.py:127.
    """
    b, c, h, w = p.shape
    cols = torch.arange(w, dtype=p.dtype)
    rows = torch.arange(h, dtype=p.dtype)
    ex = (p.sum(dim=2) * cols).sum(dim=-1)
    ey = (p.sum(dim=3) * rows).sum(dim=-1)
    return torch.stack([ex, ey], dim=-1)


_BINOMIAL5 = torch.tensor([1.0, 4.0, 6.0, 4.0, 1.0])


def pyramid_blur(x: torch.Tensor) -> torch.Tensor:
    """5x5 binomial [1,4,6,4,1]^2/256 depthwise correlation with ZERO padding.

    kornia.filters.filter2d(x, _get_pyramid_gaussian_kernel(), border_type="constant");
    call site lightning_pose/models/heads/heatmap.py:93,99.
    """
    b, c, h, w = x.shape
    k2 = torch.outer(_BINOMIAL5, _BINOMIAL5) / 256.0
    kern = k2.to(x.dtype).reshape(1, 1, 5, 5).repeat(c, 1, 1, 1)
    return F.conv2d(F.pad(x, (2, 2, 2, 2), mode="constant", value=0.0), kern, groups=c)


def kl_div_2d(pred: torch.Tensor, target: torch.Tensor) -> torch.Tensor:
    """sum_pixels target * (log target - log pred) per (B, C).

    kornia.losses.kl_div_loss_2d(pred, target, reduction="none");
    call sites lightning_pose/losses/losses.py:358,748.
    """
    b, c = pred.shape[:2]
    q = pred.reshape(b, c, -1)
    p = target.reshape(b, c, -1)
    return torch.xlogy(p, p).sub(p * torch.log(q)).sum(-1)


def js_div_2d(pred: torch.Tensor, target: torch.Tensor) -> torch.Tensor:
    """0.5*KL(target||m) + 0.5*KL(pred||m), m = (pred+target)/2.

    kornia.losses.js_div_loss_2d(reduction="none"); call site lightning_pose/losses/losses.py:402.
    """
    m = 0.5 * (pred + target)
    return 0.5 * kl_div_2d(m, target) + 0.5 * kl_div_2d(m, pred)


# --------------------------------------------------------------------------------------
# a2: heatmap head forward (lightning_pose/models/heads/heatmap.py:20-71, 203-212)
# --------------------------------------------------------------------------------------


def head_num_layers(stride: int, downsample_factor: int) -> int:
    """n_layers = log2(stride) - ds - 1 (lightning_pose/models/heads/heatmap.py:190-193)."""
    return int(math.log2(stride)) - downsample_factor - 1


def head_forward(
    features: torch.Tensor,
    weights: list[torch.Tensor],
    biases: list[torch.Tensor],
    final_softmax: bool = True,
) -> torch.Tensor:
    """PixelShuffle(2) -> ConvTranspose2d(k3,s2,p1,op1) x len(weights) -> spatial softmax(T=1).

    lightning_pose/models/heads/heatmap.py:44-71 (layer stack), :203-212 (forward).
    Weight layout (C_in, C_out, 3, 3) as torch.nn.ConvTranspose2d.
    """
    x = F.pixel_shuffle(features, 2)
    for w, b in zip(weights, biases):
        x = F.conv_transpose2d(x, w, b, stride=2, padding=1, output_padding=1)
    if final_softmax:
        x = spatial_softmax2d(x, 1.0)
    return x


def context_windows(seq: torch.Tensor, context_length: int = 5) -> torch.Tensor:
    """(n, ...) -> (n, ctx, ...): window i = padded[i : i + ctx], padded = [s0, s0, s..., s_last, s_last].

    lightning_pose/models/base.py:159-196 (get_context_from_sequence; the result there is always float32).
    """
    n = seq.shape[0]
    half = context_length // 2
    idx = (torch.arange(n)[:, None] + torch.arange(context_length)[None, :] - half).clamp(0, n - 1)
    return seq[idx].float()


def mhcrnn_multiframe(features: torch.Tensor, p: dict, upsampling_factor: int) -> torch.Tensor:
    """UpsamplingCRNN.forward: features (frames, batch, C, h, w) -> softmaxed heatmaps (batch, K, H, W).

    lightning_pose/models/heads/heatmap_mhcrnn.py:268-316.  ``p``: W_pre (w, b) [upsampling_factor 2 only], W_f, W_b
    (ConvTranspose2d k3 s2 p1 op1), H_f / H_b = (conv_w, conv_b, convt_w, convt_b) with groups = K, kernel = stride = 2.
    """
    frames, batch = features.shape[:2]
    k = p["W_f"][0].shape[1]
    ct = lambda x, wb: F.conv_transpose2d(x, wb[0], wb[1], stride=2, padding=1, output_padding=1)
    x = F.pixel_shuffle(features.reshape(frames * batch, *features.shape[2:]), 2)
    if upsampling_factor == 2:
        x = ct(x, p["W_pre"])
    x = x.reshape(frames, batch, *x.shape[1:])

    def hidden(z, hp):
        z = F.conv2d(z, hp[0], hp[1], stride=2, groups=k)
        return F.conv_transpose2d(z, hp[2], hp[3], stride=2, groups=k)

    xf = ct(x[0], p["W_f"])
    for t in range(1, frames):
        xf = ct(x[t], p["W_f"]) + hidden(xf, p["H_f"])
    xb = ct(x[frames - 1], p["W_b"])
    for t in range(frames - 2, -1, -1):
        xb = ct(x[t], p["W_b"]) + hidden(xb, p["H_b"])
    return spatial_softmax2d((xf + xb) / 2, 1.0)


# --------------------------------------------------------------------------------------
# a3/a4/a5: soft-argmax decode (heads/heatmap.py:86-144, data/heatmaps.py:90-142)
# --------------------------------------------------------------------------------------


def upsample(x: torch.Tensor) -> torch.Tensor:
    """2x bicubic (align_corners=False) then zero-padded 5x5 binomial blur.

    lightning_pose/models/heads/heatmap.py:86-100.
    """
    _, _, h, w = x.shape
    up = F.interpolate(x, size=(2 * h, 2 * w), mode="bicubic", align_corners=False)
    return pyramid_blur(up)


def confidence_window_sum(
    p: torch.Tensor, locs: torch.Tensor, sigma: float = 1.25, num_stds: int = 2
) -> torch.Tensor:
    """Sum of the (2r+1)^2 window of ``p`` around (trunc(y), trunc(x)), zero outside the plane.

    lightning_pose/data/heatmaps.py:90-142 (r = floor(sigma*num_stds) = 2; the padded copy is
    fp32 regardless of input dtype, :114-122; ``.type(int64)`` truncates toward zero, :130-131).
    """
    r = int(np.floor(sigma * num_stds))
    b, c, h, w = p.shape
    padded = torch.zeros((b, c, h + 2 * r, w + 2 * r), dtype=torch.float32)
    padded[:, :, r : r + h, r : r + w] = p
    cy = locs[..., 1].to(torch.int64) + r
    cx = locs[..., 0].to(torch.int64) + r
    offs = torch.arange(-r, r + 1)
    yy = (cy[..., None, None] + offs[:, None]).expand(b, c, 2 * r + 1, 2 * r + 1)
    xx = (cx[..., None, None] + offs[None, :]).expand(b, c, 2 * r + 1, 2 * r + 1)
    bi = torch.arange(b)[:, None, None, None]
    ci = torch.arange(c)[None, :, None, None]
    return padded[bi, ci, yy, xx].sum(dim=(-1, -2))


_DECODE_OFFSET = {1: 0.5, 2: 1.5, 3: 2.5}


def decode_softargmax(
    heatmaps: torch.Tensor, downsample_factor: int, temperature: float = 1000.0
) -> tuple[torch.Tensor, torch.Tensor]:
    """heatmaps (B,K,h,w) -> (preds (B,2K) [x0,y0,x1,y1,...], confidences (B,K)).

    lightning_pose/models/heads/heatmap.py:103-144: upsample x ds, softmax(T), expectation,
    window confidence on the pre-offset coordinates, then ``preds -= {0.5,1.5,2.5}`` (:131-136).
    """
    field = heatmaps.to(torch.float32)
    for _ in range(downsample_factor):
        field = upsample(field)
    p = spatial_softmax2d(field, temperature)
    preds = spatial_expectation2d(p)
    conf = confidence_window_sum(p, preds)
    preds = preds - _DECODE_OFFSET.get(downsample_factor, 0.0)
    return preds.reshape(-1, heatmaps.shape[1] * 2), conf


# --------------------------------------------------------------------------------------
# a6: Gaussian target generation (lightning_pose/data/heatmaps.py:11-87)
# --------------------------------------------------------------------------------------


def gaussian_targets(
    keypoints: torch.Tensor,
    height: int,
    width: int,
    output_shape: tuple[int, int],
    sigma: float = 1.25,
    visibility: torch.Tensor | None = None,
) -> torch.Tensor:
    """keypoints (B,K,2) in image pixels -> normalised Gaussian planes (B,K,oh,ow).

    lightning_pose/data/heatmaps.py:37-87.  Rules restated: scale x by ow/width, y by oh/height
    (:41-42); bad = NaN(x) | x<-1 | x>ow+1 | y<-1 | y>oh+1 after scaling (:43-49); clamp to
    [-1, size+1] (:52-53); exp(-((col-x)^2+(row-y)^2)/(2 sigma^2)) normalised by the plane sum
    (:66-72); then overwrite (:78-85): visibility None: bad->0; else 0->zeros, 1->uniform
    1/(oh*ow), (2 & bad)->zeros.
    """
    oh, ow = int(output_shape[0]), int(output_shape[1])
    kp = keypoints.detach().clone().to(torch.float32)
    x = kp[..., 0] * (ow / width)
    y = kp[..., 1] * (oh / height)
    bad = torch.isnan(x) | (x < -1) | (x > ow + 1) | (y < -1) | (y > oh + 1)
    xc = torch.clamp(x, -1, ow + 1)[..., None, None]
    yc = torch.clamp(y, -1, oh + 1)[..., None, None]
    cols = torch.arange(ow, dtype=torch.float32)[None, None, None, :]
    rows = torch.arange(oh, dtype=torch.float32)[None, None, :, None]
    g = torch.exp(-((cols - xc) ** 2 + (rows - yc) ** 2) / (2.0 * sigma**2))
    g = g / g.sum(dim=(2, 3), keepdim=True)
    zero = torch.zeros((oh, ow))
    uniform = torch.full((oh, ow), 1.0 / (oh * ow))
    if visibility is None:
        g[bad] = zero
    else:
        g[visibility == 0] = zero
        g[visibility == 1] = uniform
        g[(visibility == 2) & bad] = zero
    return g


# --------------------------------------------------------------------------------------
# a8/a9: coordinate remap (lightning_pose/data/utils.py:142-234, data/bboxes.py:74-105,222-288)
# --------------------------------------------------------------------------------------


def undo_affine(keypoints: torch.Tensor, transforms: torch.Tensor, is_multiview: bool = False):
    """Apply the inverse of the 2x3 augmentation affine to (T, 2K) keypoints.

    lightning_pose/data/utils.py:191-234: only when ``transforms.shape[-1] == 3``; a (2,3)
    transform is shared by all frames, (T,2,3) is per frame; multiview: transforms[v] applies to
    the v-th contiguous block of K/V keypoints.  Inverse = [A^-1 | -A^-1 t] (:164-167).
    """
    if transforms.shape[-1] != 3:
        return keypoints
    t_frames = keypoints.shape[0]
    kp = keypoints.reshape(t_frames, -1, 2).clone()
    k = kp.shape[1]

    def inv_apply(pts, mat):
        mat = mat.to(torch.float32)
        if mat.dim() == 2:
            mat = mat[None]
        a_inv = torch.linalg.inv(mat[:, :, :2])
        shift = -(a_inv @ mat[:, :, 2:3])
        # row-vector form used by the reference: [x y 1] @ [[A^-1]^T ; shift^T]
        return pts @ a_inv.transpose(1, 2) + shift.transpose(1, 2)

    if not is_multiview:
        kp = inv_apply(kp, transforms)
    else:
        v = transforms.shape[0]
        per = k // v
        for i in range(v):
            kp[:, i * per : (i + 1) * per] = inv_apply(kp[:, i * per : (i + 1) * per], transforms[i])
    return kp.reshape(t_frames, -1)


def model_to_frame(
    keypoints: torch.Tensor, bbox: torch.Tensor, model_height: int, model_width: int,
    num_views: int = 1,
) -> torch.Tensor:
    """x/=W_model, y/=H_model, then x*bbox_w+bbox_x, y*bbox_h+bbox_y (bbox = [x,y,h,w] per view).

    lightning_pose/data/bboxes.py:222-288 (+ norm_to_frame :74-105).  Context batches whose
    bbox has 4 more rows than the keypoints use ``bbox[2:-2]`` (:98-103).  Out of place here
    (the reference mutates its input through a view, :240-241).
    """
    n = keypoints.shape[0]
    kp = keypoints.reshape(n, -1, 2).clone().to(torch.float32)
    k = kp.shape[1]
    if bbox.shape[0] != n:
        bbox = bbox[2:-2]
    per = k // num_views
    kp[..., 0] = kp[..., 0] / model_width
    kp[..., 1] = kp[..., 1] / model_height
    for v in range(num_views):
        bb = bbox[:, 4 * v : 4 * v + 4].to(torch.float32)
        sl = slice(v * per, (v + 1) * per)
        kp[:, sl, 0] = kp[:, sl, 0] * bb[:, 3:4] + bb[:, 0:1]
        kp[:, sl, 1] = kp[:, sl, 1] * bb[:, 2:3] + bb[:, 1:2]
    return kp.reshape(n, -1)


# --------------------------------------------------------------------------------------
# a10-a16: loss stack (lightning_pose/losses/losses.py, losses/factory.py, utils/pca.py)
# --------------------------------------------------------------------------------------


def loss_weight(log_weight: float) -> float:
    """1 / (2 exp(log_weight))  (lightning_pose/losses/losses.py:89-100)."""
    return 1.0 / (2.0 * math.exp(log_weight))


def _kept_planes(targets: torch.Tensor) -> torch.Tensor:
    """planes whose target is not all-zero (lightning_pose/losses/losses.py:246-249)."""
    return ~torch.all(targets.reshape(targets.shape[0], targets.shape[1], -1) == 0.0, dim=-1)


def heatmap_mse_loss(targets: torch.Tensor, preds: torch.Tensor) -> torch.Tensor:
    """mean over kept planes' pixels of (t-p)^2 * h * w (losses.py:314-335, :246-249, :287)."""
    keep = _kept_planes(targets)
    h, w = targets.shape[-2:]
    d = (targets[keep] - preds[keep]) ** 2 * h * w
    return d.mean()


def heatmap_kl_loss(targets: torch.Tensor, preds: torch.Tensor) -> torch.Tensor:
    """mean over kept planes of KL(t+1e-10 || p+1e-10) (losses.py:360-378)."""
    keep = _kept_planes(targets)
    return kl_div_2d((preds[keep] + 1e-10)[None], (targets[keep] + 1e-10)[None])[0].mean()


def heatmap_js_loss(targets: torch.Tensor, preds: torch.Tensor) -> torch.Tensor:
    """mean over kept planes of JS(t+1e-10, p+1e-10) (losses.py:404-423)."""
    keep = _kept_planes(targets)
    return js_div_2d((preds[keep] + 1e-10)[None], (targets[keep] + 1e-10)[None])[0].mean()


def _confidence_pair_mask(conf: torch.Tensor, thr: float) -> torch.Tensor:
    low = conf < thr
    return low[:-1] | low[1:]


def temporal_loss(
    keypoints: torch.Tensor,
    confidences: torch.Tensor | None = None,
    epsilon: float | list[float] = 0.0,
    prob_threshold: float = 0.0,
) -> torch.Tensor:
    """mean over (T-1, K) of relu(||kp[t+1]-kp[t]||_2 masked by confidence - eps_k).

    lightning_pose/losses/losses.py:608-703: norms (:651-672), zero where conf[t] or conf[t+1]
    < threshold BEFORE the epsilon rectification (:622-649), per-keypoint epsilon (:608-620),
    mean over all entries including the masked ones (:700-701).
    """
    t = keypoints.shape[0]
    d = torch.diff(keypoints, dim=0).reshape(t - 1, -1, 2)
    n = torch.linalg.norm(d, ord=2, dim=2)
    if confidences is not None:
        n = n.masked_fill(_confidence_pair_mask(confidences, prob_threshold), 0.0)
    eps = torch.as_tensor(epsilon, dtype=torch.float32)
    return F.relu(n - eps).mean()


def temporal_heatmap_loss(
    heatmaps: torch.Tensor,
    confidences: torch.Tensor,
    kind: str = "mse",
    epsilon: float | list[float] = 0.0,
    prob_threshold: float = 0.0,
) -> torch.Tensor:
    """per consecutive pair: mean-pixel MSE or KL(h[t]+1e-10 as pred, h[t+1]+1e-10 as target).

    lightning_pose/losses/losses.py:793-854 (argument order of the KL call at :818-822:
    ``hmloss(pred=predictions[i], target=predictions[i+1])``).
    """
    a, b = heatmaps[:-1], heatmaps[1:]
    if kind == "mse":
        d = ((a - b) ** 2).reshape(a.shape[0], a.shape[1], -1).mean(-1)
    elif kind == "kl":
        d = kl_div_2d(a + 1e-10, b + 1e-10)
    else:
        raise ValueError(kind)
    d = d.masked_fill(_confidence_pair_mask(confidences, prob_threshold), 0.0)
    eps = torch.as_tensor(epsilon, dtype=torch.float32)
    return F.relu(d - eps).mean()


def pca_format_singleview(
    keypoints: torch.Tensor, columns: list[int] | None = None, centering: str | None = None
) -> torch.Tensor:
    """(T,2K) -> (T,2K_sel): select keypoints, optional centroid removal (utils/pca.py:124-163)."""
    t = keypoints.shape[0]
    kp = keypoints.reshape(t, -1, 2)
    if columns is not None:
        kp = kp[:, np.asarray(columns), :]
    if centering == "mean":
        kp = kp - kp.mean(dim=1, keepdim=True)
    elif centering == "median":
        kp = kp - kp.quantile(dim=1, q=0.5, keepdim=True)
    elif centering is not None:
        raise NotImplementedError(centering)
    return kp.reshape(t, -1)


def pca_format_multiview(keypoints: torch.Tensor, mirrored_column_matches) -> torch.Tensor:
    """(T,2K) -> (T*K_mv, 2V); row index = t*K_mv + j, columns [x_v0,y_v0,x_v1,y_v1,...].

    lightning_pose/utils/pca.py:97-122 and :759-792 (``permute(2,0,1).reshape(2,-1)`` per view,
    concatenated over views, transposed).
    """
    t = keypoints.shape[0]
    kp = keypoints.reshape(t, -1, 2)
    cols = []
    for view_cols in mirrored_column_matches:
        sel = kp[:, np.asarray(view_cols), :]  # (T, K_mv, 2)
        cols.append(sel.reshape(-1, 2))  # rows ordered (t, j)
    return torch.cat(cols, dim=1)


def pca_reprojection_error(data: torch.Tensor, mean: torch.Tensor, kept: torch.Tensor):
    """|| x - (((x-mu) V^T) V + mu) ||_2 per (x,y) pair -> (N, D/2)  (utils/pca.py:266-309)."""
    centered = data - mean[None]
    reproj = (centered @ kept.T) @ kept + mean[None]
    diff = (data - reproj).reshape(data.shape[0], data.shape[1] // 2, 2)
    return torch.linalg.norm(diff, dim=2)


def pca_loss(data_formatted, mean, kept, epsilon) -> torch.Tensor:
    """mean(relu(reprojection_error - eps))  (lightning_pose/losses/losses.py:548-573)."""
    err = pca_reprojection_error(data_formatted, mean, kept)
    return F.relu(err - torch.as_tensor(epsilon, dtype=torch.float32)).mean()


def reprojection_heatmap_loss(
    heatmaps_targ: torch.Tensor, keypoints_2d: torch.Tensor, height: int, width: int,
    output_shape: tuple[int, int],
) -> torch.Tensor:
    """generate_heatmaps(kp) vs target, squared error over planes with a non-zero target.

    lightning_pose/losses/losses.py:1177-1260.  Reference quirk, replicated: ``compute_loss``
    (:1216-1219) is handed the 4-D (B,K,h,w) tensors and scales by ``shape[1]*shape[2]`` =
    K*h (not h*w as HeatmapMSELoss does on its 3-D gathered planes).
    """
    pred = gaussian_targets(keypoints_2d, height, width, output_shape)
    keep = _kept_planes(heatmaps_targ)
    scale = heatmaps_targ.shape[1] * heatmaps_targ.shape[2]
    return ((heatmaps_targ[keep] - pred[keep]) ** 2 * scale).mean()


def combine_losses(named_losses: dict[str, tuple[torch.Tensor, float]], anneal_weight: float = 1.0):
    """sum_l anneal_l * weight_l * loss_l; anneal_l = 1 for heatmap_{mse,kl,js}.

    lightning_pose/losses/factory.py:229-285 (:267 exemption).
    ``named_losses[name] = (scalar_loss, log_weight)``.
    """
    total = torch.tensor(0.0)
    for name, (val, log_w) in named_losses.items():
        a = 1.0 if name in ("heatmap_mse", "heatmap_kl", "heatmap_js") else anneal_weight
        total = total + a * loss_weight(log_w) * val
    return total


# --------------------------------------------------------------------------------------
# explicit 1-D form of the decode's upsampling operator (used to pin the CUDA tap tables)
# --------------------------------------------------------------------------------------


def _bicubic_coeffs(t: float, a: float = -0.75) -> list[float]:
    """PyTorch cubic convolution coefficients (aten UpSample.h get_cubic_upsample_coefficients)."""

    def c1(x):  # |x| <= 1
        return ((a + 2.0) * x - (a + 3.0)) * x * x + 1.0

    def c2(x):  # 1 < |x| < 2
        return ((a * x - 5.0 * a) * x + 8.0 * a) * x - 4.0 * a

    return [c2(t + 1.0), c1(t), c1(1.0 - t), c2(2.0 - t)]


def upsample_matrix_1d(n: int, stages: int) -> np.ndarray:
    """Dense (n*2^stages, n) float64 matrix U with upsample^stages(h) == U_H h U_W^T.

    One stage = zero-padded [1,4,6,4,1]/16 blur applied after 2x bicubic with
    align_corners=False and edge-clamped taps (lightning_pose/models/heads/heatmap.py:86-100;
    SURVEY Appendix A.1).
    """
    total = np.eye(n)
    cur = n
    for _ in range(stages):
        bic = np.zeros((2 * cur, cur))
        for i in range(2 * cur):
            src = (i + 0.5) / 2.0 - 0.5
            i0 = math.floor(src)
            w = _bicubic_coeffs(src - i0)
            for k in range(4):
                bic[i, min(max(i0 - 1 + k, 0), cur - 1)] += w[k]
        blur = np.zeros((2 * cur, 2 * cur))
        for i in range(2 * cur):
            for k, bw in enumerate([1, 4, 6, 4, 1]):
                j = i - 2 + k
                if 0 <= j < 2 * cur:
                    blur[i, j] += bw / 16.0
        total = blur @ bic @ total
        cur *= 2
    return total
