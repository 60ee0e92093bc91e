"""torch-facing wrappers of the lpb200 C-ABI: custom ops (``torch.library``) + autograd.

PyTorch is plumbing here (device memory, the current CUDA stream, autograd bookkeeping); every
numerical operation below runs in ``liblpb200.so``.  All wrappers refuse non-CUDA tensors: there is
no CPU fallback.
"""
from __future__ import annotations

import ctypes as C

import torch

from ._lib import PcaDesc, check, lib

__all__ = [
    "decode_softargmax",
    "upsample2x",
    "generate_heatmaps",
    "keypoints_mask_oob",
    "evaluate_heatmaps_at_location",
    "head_forward",
    "head_forward_f32",
    "head_bf16_supported",
    "convt_forward_f32",
    "convt_backward_f32",
    "head_backward_bf16",
    "decode_backward_windows",
    "remap_keypoints",
    "heatmap_loss",
    "heatmap_mse_from_keypoints",
    "temporal_heatmap_loss",
    "unsup_losses",
    "PcaParams",
]

_KIND = {"mse": 0, "kl": 1, "js": 2}


def _stream() -> C.c_void_p:
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _ptr(t: torch.Tensor | None) -> C.c_void_p:
    return C.c_void_p(0 if t is None else t.data_ptr())


def _cuda_f32(t: torch.Tensor, name: str) -> torch.Tensor:
    if not isinstance(t, torch.Tensor) or not t.is_cuda:
        raise RuntimeError(f"lpb200: `{name}` must be a CUDA tensor (this package has no CPU fallback)")
    if t.dtype != torch.float32:
        t = t.float()
    return t.contiguous()


# =====================================================================================
# soft-argmax decode
# =====================================================================================
@torch.library.custom_op("lpb200::decode_fwd", mutates_args=())
def _decode_fwd(heatmaps: torch.Tensor, ds: int, temperature: float) -> tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
    b, k, h, w = heatmaps.shape
    xy = torch.empty((b, k, 2), device=heatmaps.device, dtype=torch.float32)
    conf = torch.empty((b, k), device=heatmaps.device, dtype=torch.float32)
    stats = torch.empty((b, k, 8), device=heatmaps.device, dtype=torch.float32)
    with torch.cuda.device(heatmaps.device):
        check(lib.lpb_decode_fwd(_ptr(heatmaps), b * k, h, w, ds, temperature, _ptr(xy), _ptr(conf), _ptr(stats), _stream()))
    return xy, conf, stats


def decode_forward_hinted(heatmaps: torch.Tensor, ds: int, temperature: float, hints):
    """``_decode_fwd`` with the per-plane hints the bf16 head's softmax pass wrote for these very heatmaps (or None):
    peaked planes are decoded from the window around their maximum without a sweep of the plane; same results.
    No autograd node of its own (``_HeadFunction`` owns the backward)."""
    b, k, h, w = heatmaps.shape
    xy = torch.empty((b, k, 2), device=heatmaps.device, dtype=torch.float32)
    conf = torch.empty((b, k), device=heatmaps.device, dtype=torch.float32)
    stats = torch.empty((b, k, 8), device=heatmaps.device, dtype=torch.float32)
    with torch.cuda.device(heatmaps.device):
        check(lib.lpb_decode_fwd_hinted(_ptr(heatmaps), b * k, h, w, ds, temperature, _ptr(xy), _ptr(conf), _ptr(stats), _ptr(hints), _stream()))
    return xy, conf, stats


@_decode_fwd.register_fake
def _(heatmaps, ds, temperature):
    b, k, _, _ = heatmaps.shape
    return heatmaps.new_empty((b, k, 2)), heatmaps.new_empty((b, k)), heatmaps.new_empty((b, k, 8))


@torch.library.custom_op("lpb200::decode_bwd", mutates_args=())
def _decode_bwd(heatmaps: torch.Tensor, stats: torch.Tensor, grad_xy: torch.Tensor, ds: int, temperature: float) -> torch.Tensor:
    b, k, h, w = heatmaps.shape
    g = torch.empty_like(heatmaps)
    with torch.cuda.device(heatmaps.device):
        check(lib.lpb_decode_bwd(_ptr(heatmaps), _ptr(stats), _ptr(grad_xy), b * k, h, w, ds, temperature, _ptr(g), _stream()))
    return g


@_decode_bwd.register_fake
def _(heatmaps, stats, grad_xy, ds, temperature):
    return torch.empty_like(heatmaps)


def _decode_setup(ctx, inputs, output):
    heatmaps, ds, temperature = inputs
    ctx.save_for_backward(heatmaps, output[2])
    ctx.ds, ctx.temperature = ds, temperature


def _decode_backward(ctx, g_xy, g_conf, g_stats):
    heatmaps, stats = ctx.saved_tensors
    if g_xy is None:
        return None, None, None
    return _decode_bwd(heatmaps, stats, g_xy.contiguous().float(), ctx.ds, ctx.temperature), None, None


_decode_fwd.register_autograd(_decode_backward, setup_context=_decode_setup)


def decode_softargmax(heatmaps: torch.Tensor, downsample_factor: int, temperature: float = 1000.0):
    """(B,K,h,w) heatmaps -> (preds (B,2K), confidences (B,K)); differentiable wrt heatmaps (via preds)."""
    hm = _cuda_f32(heatmaps, "heatmaps")
    if hm.dim() != 4:
        raise ValueError(f"heatmaps must be (batch, keypoints, h, w); got {tuple(hm.shape)}")
    if downsample_factor not in (1, 2, 3):
        raise ValueError(f"downsample_factor must be 1, 2 or 3; got {downsample_factor}")
    xy, conf, _ = _decode_fwd(hm, int(downsample_factor), float(temperature))
    return xy.reshape(-1, hm.shape[1] * 2), conf


def upsample2x(inputs: torch.Tensor) -> torch.Tensor:
    x = _cuda_f32(inputs, "inputs")
    b, k, h, w = x.shape
    out = torch.empty((b, k, 2 * h, 2 * w), device=x.device, dtype=torch.float32)
    with torch.cuda.device(x.device):
        check(lib.lpb_upsample2x(_ptr(x), b * k, h, w, _ptr(out), _stream()))
    return out


# =====================================================================================
# Gaussian targets / windowed evaluation
# =====================================================================================
class _GenerateHeatmaps(torch.autograd.Function):
    @staticmethod
    def forward(ctx, keypoints, visibility, height, width, oh, ow, sigma):
        kp = keypoints.contiguous().float()
        b, k, _ = kp.shape
        out = torch.empty((b, k, oh, ow), device=kp.device, dtype=torch.float32)
        with torch.cuda.device(kp.device):
            check(lib.lpb_generate_heatmaps(_ptr(kp), _ptr(visibility), b * k, float(height), float(width), oh, ow, float(sigma), _ptr(out), _stream()))
        ctx.save_for_backward(kp, visibility)
        ctx.meta = (float(height), float(width), oh, ow, float(sigma))
        return out

    @staticmethod
    def backward(ctx, g):
        kp, vis = ctx.saved_tensors
        height, width, oh, ow, sigma = ctx.meta
        gk = torch.empty_like(kp)
        with torch.cuda.device(kp.device):
            check(lib.lpb_generate_heatmaps_bwd(_ptr(kp), _ptr(vis), _ptr(g.contiguous().float()), kp.shape[0] * kp.shape[1], height, width, oh, ow, sigma, _ptr(gk), _stream()))
        return gk, None, None, None, None, None, None


def generate_heatmaps(keypoints, height, width, output_shape, sigma=1.25, keep_gradients=False, visibility=None):
    kp = _cuda_f32(keypoints, "keypoints")
    if kp.dim() != 3 or kp.shape[-1] != 2:
        raise ValueError(f"keypoints must be (batch, num_keypoints, 2); got {tuple(kp.shape)}")
    vis = None
    if visibility is not None:
        if not visibility.is_cuda:
            raise RuntimeError("lpb200: `visibility` must be a CUDA tensor")
        vis = visibility.to(torch.int32).contiguous()
    if not keep_gradients:
        kp = kp.detach()
    return _GenerateHeatmaps.apply(kp, vis, height, width, int(output_shape[0]), int(output_shape[1]), sigma)


def keypoints_mask_oob(keypoints, height, width):
    """Keypoints outside [0, width) x [0, height) -> NaN in both coordinates (labeled-data rule, datasets.py:496-508)."""
    kp = _cuda_f32(keypoints, "keypoints")
    out = torch.empty_like(kp)
    with torch.cuda.device(kp.device):
        check(lib.lpb_keypoints_mask_oob(_ptr(kp), kp.numel() // 2, float(height), float(width), _ptr(out), _stream()))
    return out


def evaluate_heatmaps_at_location(heatmaps, locs, radius: int = 2):
    hm = _cuda_f32(heatmaps, "heatmaps")
    lc = _cuda_f32(locs, "locs")
    b, k, h, w = hm.shape
    out = torch.empty((b, k), device=hm.device, dtype=torch.float32)
    with torch.cuda.device(hm.device):
        check(lib.lpb_evaluate_heatmaps_at_location(_ptr(hm), _ptr(lc), b * k, h, w, int(radius), _ptr(out), _stream()))
    return out


# =====================================================================================
# heatmap head
# =====================================================================================
_BWD_WIDTHS = (4, 8, 12, 16, 24, 32)


def head_bf16_supported(shape, channels, train: bool) -> bool:
    """Whether the tcgen05 kernels cover this head: feature shape (B, C, H, W), deconv output channels (c1[, c2])."""
    _, c, h, w = shape
    n = len(channels)
    if n not in (1, 2) or c % 128 or (h * w) % 8:
        return False
    if (n == 1 and channels[0] > 20) or (n == 2 and (channels[0] >= 20 or channels[1] > 20)):
        return False
    plan = C.c_int(0)
    if lib.lpb_head_bf16_plan(c, h, w, channels[0], channels[1] if n == 2 else 0, C.byref(plan)) != 0:
        return False
    if train and (h % 2 or w not in _BWD_WIDTHS or (304 // (2 * w + 1)) < 4):
        return False
    return True


def _head_forward_bf16(f, weights, biases, final_softmax, train=False, want_hints=False):
    """tcgen05 path (one- or two-deconv heads); the caller checks ``head_bf16_supported`` first.

    ``train=True`` returns ``(out, saved)`` where ``saved`` carries what ``head_backward_bf16`` needs
    (the row-layout copy of the shuffled features and the forward workspace with the inter-layer activations).
    """
    b, c, h, w = f.shape
    n = len(weights)
    w1 = _cuda_f32(weights[0], "weight")
    b1 = _cuda_f32(biases[0], "bias")
    w2 = _cuda_f32(weights[1], "weight") if n == 2 else None
    b2 = _cuda_f32(biases[1], "bias") if n == 2 else None
    c1, c2 = w1.shape[1], (w2.shape[1] if n == 2 else 0)
    plan = C.c_int(0)
    check(lib.lpb_head_bf16_plan(c, h, w, c1, c2, C.byref(plan)))
    nbytes = C.c_size_t(0)
    check(lib.lpb_head_bf16_workspace_bytes(b, c, h, w, c1, c2, C.byref(nbytes)))
    ws = torch.empty((nbytes.value,), device=f.device, dtype=torch.uint8)
    up = 8 if n == 2 else 4
    out = torch.empty((b, c2 if n == 2 else c1, up * h, up * w), device=f.device, dtype=torch.float32)
    xs = None
    if train or plan.value == 0:  # row-layout copy of the shuffled features: the banded path's operand / the wgrad's input
        check(lib.lpb_head_bf16_saved_bytes(b, c, h, w, C.byref(nbytes)))
        xs = torch.empty((nbytes.value,), device=f.device, dtype=torch.uint8)
    # decode hints (want_hints: a soft-argmax decode of `out` follows): 16 bytes per plane, see lpb_decode_fwd_hinted
    hints = None
    if want_hints and final_softmax and lib.lpb_get_tuning(14) == 1:  # LPB_TUNE_DECODE_HINTS
        hints = torch.empty((b * out.shape[1], 4), device=f.device, dtype=torch.int32)
    with torch.cuda.device(f.device):
        check(lib.lpb_head_fwd_bf16_hinted(_ptr(f), b, c, h, w, _ptr(w1), _ptr(b1), c1, _ptr(w2), _ptr(b2), c2, int(bool(final_softmax)), _ptr(out), _ptr(xs), _ptr(ws), _ptr(hints), _stream()))
    if want_hints:
        return (out, (xs, ws), hints) if train else (out, hints)
    if train:
        return out, (xs, ws)
    return out


def decode_backward_windows(heatmaps, stats, grad_xy, ds, temperature):
    """Sparse soft-argmax gradient for the fused head backward: (win, meta, overflow) of ``lpb_decode_bwd_windows``."""
    b, k, h, w = heatmaps.shape
    win = torch.empty((b * k, 32, 32), device=heatmaps.device, dtype=torch.float32)
    meta = torch.empty((b * k, 4), device=heatmaps.device, dtype=torch.int32)
    # only planes whose support overflows a window are ever written/read here (none for peaked heatmaps);
    # the caching allocator hands the block back without touching it
    overflow = torch.empty_like(heatmaps)
    queue = torch.empty((b * k + 1,), device=heatmaps.device, dtype=torch.int32)
    with torch.cuda.device(heatmaps.device):
        check(lib.lpb_decode_bwd_windows(_ptr(heatmaps), _ptr(stats), _ptr(grad_xy), b * k, h, w, ds, temperature, _ptr(win), _ptr(meta), _ptr(overflow), _ptr(queue), _stream()))
    return win, meta, overflow


def head_backward_bf16(g_out, saved, feat_shape, w1, w2, need_dfeat=True, probs=None, windows=None):
    """Gradients of the bf16 head (tcgen05): returns (dfeat bf16 | None, dw1, db1, dw2 | None, db2 | None).

    ``g_out``: dense gradient w.r.t. the head output or None; ``probs``: the head output when it ends in the
    spatial softmax (its backward is fused in), None for a logits head; ``windows``: result of
    ``decode_backward_windows`` or None; ``w2`` None for a one-deconv head.
    """
    xs, fws = saved
    b, c, h, w = feat_shape
    g = _cuda_f32(g_out, "g_out") if g_out is not None else None
    if g is None and windows is None:
        raise ValueError("head_backward_bf16 needs a dense gradient and/or decode windows")
    w1 = _cuda_f32(w1, "w1")
    w2 = _cuda_f32(w2, "w2") if w2 is not None else None
    c1, c2 = w1.shape[1], (w2.shape[1] if w2 is not None else 0)
    dev = w1.device
    nbytes = C.c_size_t(0)
    check(lib.lpb_head_bwd_bf16_workspace_bytes(b, c, h, w, c1, c2, C.byref(nbytes)))
    ws = torch.empty((nbytes.value,), device=dev, dtype=torch.uint8)
    dfeat = torch.empty((b, c, h, w), device=dev, dtype=torch.bfloat16) if need_dfeat else None
    dw1 = torch.empty_like(w1)
    db1 = torch.empty((c1,), device=dev, dtype=torch.float32)
    dw2 = torch.empty_like(w2) if w2 is not None else None
    db2 = torch.empty((c2,), device=dev, dtype=torch.float32) if w2 is not None else None
    win, meta, ov = windows if windows is not None else (None, None, None)
    with torch.cuda.device(dev):
        check(lib.lpb_head_bwd_bf16(_ptr(g), _ptr(probs), _ptr(win), _ptr(meta), _ptr(ov), _ptr(xs), _ptr(fws), b, c, h, w, _ptr(w1), c1, _ptr(w2), c2,
                                    _ptr(dfeat), _ptr(dw1), _ptr(db1), _ptr(dw2), _ptr(db2), _ptr(ws), _stream()))
    return dfeat, dw1, db1, dw2, db2


# ---- fp32 precision path: one transposed convolution at a time (any number of layers), native backward ----------
def convt_forward_f32(x, weight, bias, shuffle: bool):
    """[PixelShuffle(2) +] ConvTranspose2d(k3, s2, p1, op1) in fp32 on CUDA cores."""
    x = _cuda_f32(x, "input")
    wt = _cuda_f32(weight, "weight")
    bs = _cuda_f32(bias, "bias") if bias is not None else None
    b, c, h, w = x.shape
    cin, hi, wi = (c // 4, 2 * h, 2 * w) if shuffle else (c, h, w)
    if wt.shape[0] != cin:
        raise ValueError(f"weight expects {wt.shape[0]} input channels, got {cin}")
    out = torch.empty((b, wt.shape[1], 2 * hi, 2 * wi), device=x.device, dtype=torch.float32)
    with torch.cuda.device(x.device):
        check(lib.lpb_convt_fwd_f32(_ptr(x), b, cin, hi, wi, int(shuffle), _ptr(wt), _ptr(bs), wt.shape[1], _ptr(out), _stream()))
    return out


def convt_backward_f32(x, grad_out, weight, shuffle: bool, need_dx: bool = True, need_db: bool = True):
    """Autograd of ``convt_forward_f32``: (dx | None, dw, db | None)."""
    x = _cuda_f32(x, "input")
    g = _cuda_f32(grad_out, "grad_out")
    wt = _cuda_f32(weight, "weight")
    b, c, h, w = x.shape
    cin, hi, wi = (c // 4, 2 * h, 2 * w) if shuffle else (c, h, w)
    dx = torch.empty_like(x) if need_dx else None
    dw = torch.empty_like(wt)
    db = torch.empty((wt.shape[1],), device=x.device, dtype=torch.float32) if need_db else None
    with torch.cuda.device(x.device):
        check(lib.lpb_convt_bwd_f32(_ptr(x), _ptr(g), b, cin, hi, wi, int(shuffle), _ptr(wt), wt.shape[1], _ptr(dx), _ptr(dw), _ptr(db), _stream()))
    return dx, dw, db


def plane_softmax_(x: torch.Tensor) -> torch.Tensor:
    """In-place spatial softmax (T = 1) over each (b, k) plane of a contiguous fp32 tensor."""
    b, k, h, w = x.shape
    with torch.cuda.device(x.device):
        check(lib.lpb_plane_softmax_f32(_ptr(x), b * k, h * w, _stream()))
    return x


def head_forward_f32(features, weights, biases, final_softmax=True, keep_activations=False):
    """fp32 head: PixelShuffle(2) + any number of deconvs + spatial softmax.  Returns ``out`` or, with
    ``keep_activations``, ``(out, inputs_of_each_layer)`` for ``convt_backward_f32``."""
    x = _cuda_f32(features, "features")
    acts = []
    for i, (wt, bs) in enumerate(zip(weights, biases)):
        acts.append(x)
        x = convt_forward_f32(x, wt, bs, shuffle=(i == 0))
    if final_softmax:
        plane_softmax_(x)
    return (x, acts) if keep_activations else x


def head_forward(features, weights, biases, final_softmax=True):
    """PixelShuffle(2) + ConvTranspose2d stack + spatial softmax; forward only.

    bf16 features take the tcgen05 tensor-core kernels (fp32 accumulate, fp32 heatmaps) when the head is a one- or
    two-deconv head the tiling covers; everything else takes the full-precision CUDA-core kernels.  Training uses
    ``HeatmapHead`` which adds the backward.
    """
    if not isinstance(features, torch.Tensor) or not features.is_cuda:
        raise RuntimeError("lpb200: `features` must be a CUDA tensor (this package has no CPU fallback)")
    if len(weights) < 1:
        raise ValueError("head needs at least one deconv layer")
    if features.dtype == torch.bfloat16 and all(b is not None for b in biases) and head_bf16_supported(
            tuple(features.shape), [w.shape[1] for w in weights], train=False):
        return _head_forward_bf16(features.contiguous(), weights, biases, final_softmax)
    return head_forward_f32(features, weights, biases, final_softmax)


# =====================================================================================
# coordinate remap
# =====================================================================================
class _Remap(torch.autograd.Function):
    @staticmethod
    def forward(ctx, kp, tf, per_frame, num_views, bb, model_height, model_width, out):
        n, k2 = kp.shape
        res = out if out is not None else torch.empty_like(kp)
        with torch.cuda.device(kp.device):
            check(lib.lpb_remap_keypoints(_ptr(kp), n, k2 // 2, _ptr(tf), per_frame, num_views, _ptr(bb), bb.shape[0], model_height, model_width, _ptr(res), _stream()))
        ctx.save_for_backward(tf, bb)
        ctx.meta = (per_frame, num_views, model_height, model_width)
        if out is not None:
            ctx.mark_dirty(out)
        return res

    @staticmethod
    def backward(ctx, g):
        tf, bb = ctx.saved_tensors
        per_frame, num_views, mh, mw = ctx.meta
        g = g.contiguous().float()
        n, k2 = g.shape
        gi = torch.empty_like(g)
        with torch.cuda.device(g.device):
            check(lib.lpb_remap_keypoints_bwd(_ptr(g), n, k2 // 2, _ptr(tf), per_frame, num_views, _ptr(bb), bb.shape[0], mh, mw, _ptr(gi), _stream()))
        return gi, None, None, None, None, None, None, None


def remap_keypoints(keypoints, transforms, bbox, model_height, model_width, is_multiview=False, num_views=1, out=None):
    """undo_affine_transform_batch + model_to_frame_batch in one launch; differentiable in ``keypoints``."""
    kp = _cuda_f32(keypoints, "keypoints")
    n, k2 = kp.shape
    tf = None
    per_frame = 0
    if transforms is not None and transforms.shape[-1] == 3:
        tf = _cuda_f32(transforms, "transforms")
        if not is_multiview and tf.dim() == 3:
            if tf.shape[0] == 1:  # a lone (1, 2, 3) transform is tiled over the frames (data/utils.py:193-200)
                tf = tf[0].contiguous()
            elif tf.shape[0] == n:
                per_frame = 1
            else:
                raise ValueError(f"per-frame transforms {tuple(tf.shape)} vs {n} frames")
    bb = _cuda_f32(bbox, "bbox")
    if out is not None and kp.requires_grad:
        out = None  # writing through a tensor that autograd tracks is not allowed; return a fresh one
    return _Remap.apply(kp, tf, per_frame, int(num_views), bb, float(model_height), float(model_width), out)


# =====================================================================================
# MHCRNN context branch
# =====================================================================================
class _CrnnCombine(torch.autograd.Function):
    """(x_f + x_b) / 2 of the bidirectional conv-RNN, from per-frame deconv maps and a window index table."""

    @staticmethod
    def forward(ctx, wf, wb, idx, cwf, cbf, twf, tbf, cwb, cbb, twb, tbb):
        n, k, h, w = wf.shape
        m = idx.shape[0]
        f = cwf.shape[0] // k
        dev = wf.device
        Lf, hf = torch.empty((k, 16), device=dev), torch.empty((k, 4), device=dev)
        Lb, hb = torch.empty((k, 16), device=dev), torch.empty((k, 4), device=dev)
        out = torch.empty((m, k, h, w), device=dev, dtype=torch.float32)
        with torch.cuda.device(dev):
            check(lib.lpb_crnn_prepare(_ptr(cwf), _ptr(cbf), _ptr(twf), _ptr(tbf), k, f, _ptr(Lf), _ptr(hf), _stream()))
            check(lib.lpb_crnn_prepare(_ptr(cwb), _ptr(cbb), _ptr(twb), _ptr(tbb), k, f, _ptr(Lb), _ptr(hb), _stream()))
            check(lib.lpb_crnn_combine_fwd(_ptr(wf), _ptr(wb), _ptr(idx), m, n, k, h, w, _ptr(Lf), _ptr(hf), _ptr(Lb), _ptr(hb), _ptr(out), _stream()))
        ctx.save_for_backward(wf, wb, idx, cwf, cbf, twf, cwb, cbb, twb, Lf, hf, Lb, hb)
        return out

    @staticmethod
    def backward(ctx, g):
        wf, wb, idx, cwf, cbf, twf, cwb, cbb, twb, Lf, hf, Lb, hb = ctx.saved_tensors
        n, k, h, w = wf.shape
        m = idx.shape[0]
        f = cwf.shape[0] // k
        dev = wf.device
        g = g.contiguous().float()
        dwf, dwb = torch.empty_like(wf), torch.empty_like(wb)
        dLf, dhf, dLb, dhb = torch.empty_like(Lf), torch.empty_like(hf), torch.empty_like(Lb), torch.empty_like(hb)
        outs = [torch.empty_like(t) for t in (cwf, cbf, twf)] + [torch.empty((k,), device=dev)] + [torch.empty_like(t) for t in (cwb, cbb, twb)] + [torch.empty((k,), device=dev)]
        with torch.cuda.device(dev):
            check(lib.lpb_crnn_combine_bwd(_ptr(wf), _ptr(wb), _ptr(idx), _ptr(g), m, n, k, h, w, _ptr(Lf), _ptr(hf), _ptr(Lb), _ptr(hb),
                                           _ptr(dwf), _ptr(dwb), _ptr(dLf), _ptr(dhf), _ptr(dLb), _ptr(dhb), _stream()))
            check(lib.lpb_crnn_prepare_bwd(_ptr(cwf), _ptr(cbf), _ptr(twf), _ptr(dLf), _ptr(dhf), k, f, _ptr(outs[0]), _ptr(outs[1]), _ptr(outs[2]), _ptr(outs[3]), _stream()))
            check(lib.lpb_crnn_prepare_bwd(_ptr(cwb), _ptr(cbb), _ptr(twb), _ptr(dLb), _ptr(dhb), k, f, _ptr(outs[4]), _ptr(outs[5]), _ptr(outs[6]), _ptr(outs[7]), _stream()))
        return (dwf, dwb, None, *outs)


def crnn_combine(wf, wb, idx, h_f_params, h_b_params):
    """Pre-softmax MHCRNN logits (M, K, H, W).  ``wf`` / ``wb``: (N, K, H, W) maps W_f(x_t) / W_b(x_t) of N frames;
    ``idx``: (M, 5) int32 frame indices of the context slots; ``h_*_params`` = (conv.weight, conv.bias, convT.weight,
    convT.bias) of ``H_f`` / ``H_b``.  Differentiable in the maps and in the eight parameter tensors."""
    wf, wb = _cuda_f32(wf, "wf"), _cuda_f32(wb, "wb")
    if wf.shape != wb.shape or wf.dim() != 4 or wf.shape[2] % 2 or wf.shape[3] % 2:
        raise ValueError(f"deconv maps must be (N, K, H, W) with even H, W; got {tuple(wf.shape)} / {tuple(wb.shape)}")
    idx = idx.to(device=wf.device, dtype=torch.int32).contiguous()
    if idx.dim() != 2 or idx.shape[1] != 5:
        raise ValueError(f"idx must be (M, 5); got {tuple(idx.shape)}")
    ps = [_cuda_f32(t, "crnn parameter") for t in (*h_f_params, *h_b_params)]
    k = wf.shape[1]
    if idx.shape[0] * k >= 65536:  # grid.y limit of one launch: split the windows
        parts = [_CrnnCombine.apply(wf, wb, idx[i : i + 65535 // k], *ps) for i in range(0, idx.shape[0], 65535 // k)]
        return torch.cat(parts, dim=0)
    return _CrnnCombine.apply(wf, wb, idx, *ps)


class _PlaneSoftmax(torch.autograd.Function):
    @staticmethod
    def forward(ctx, logits):
        p = logits.contiguous().float().clone()
        plane_softmax_(p)
        ctx.save_for_backward(p)
        return p

    @staticmethod
    def backward(ctx, g):
        (p,) = ctx.saved_tensors
        return plane_softmax_backward(p, g)


def plane_softmax(logits: torch.Tensor) -> torch.Tensor:
    """spatial_softmax2d(x, temperature=1) with its native backward."""
    return _PlaneSoftmax.apply(_cuda_f32(logits, "logits"))


def context_gather(seq: torch.Tensor, context_length: int = 5) -> torch.Tensor:
    """get_context_from_sequence (base.py:159-196): (n, ...) -> (n, ctx, ...) windows with replicated edges."""
    if not isinstance(seq, torch.Tensor) or not seq.is_cuda:
        raise RuntimeError("lpb200: `img_seq` must be a CUDA tensor (this package has no CPU fallback)")
    x = seq.contiguous()
    n = x.shape[0]
    item = x[0].numel() * x.element_size()
    if item % 16:
        raise ValueError(f"items of {item} bytes: the gather moves 16-byte vectors")
    out = torch.empty((n, context_length, *x.shape[1:]), device=x.device, dtype=x.dtype)
    with torch.cuda.device(x.device):
        check(lib.lpb_context_gather(_ptr(x), n, item, int(context_length), _ptr(out), _stream()))
    return out


IMAGENET_MEAN = (0.485, 0.456, 0.406)  # lightning_pose/data/__init__.py (_IMAGENET_MEAN / _IMAGENET_STD)
IMAGENET_STD = (0.229, 0.224, 0.225)


def frames_normalize(frames_u8, size=None, mean=IMAGENET_MEAN, std=IMAGENET_STD, channels_last=False, dtype=torch.float32):
    """uint8 (F, H, W, 3) decoded RGB frames -> normalised (F, 3, h, w) [or (F, h, w, 3)] fp32 / bf16 in one pass."""
    if not isinstance(frames_u8, torch.Tensor) or not frames_u8.is_cuda:
        raise RuntimeError("lpb200: `frames` must be a CUDA tensor (this package has no CPU fallback)")
    if frames_u8.dtype != torch.uint8 or frames_u8.dim() != 4 or frames_u8.shape[-1] != 3:
        raise ValueError(f"frames must be uint8 (F, H, W, 3); got {tuple(frames_u8.shape)} {frames_u8.dtype}")
    if dtype not in (torch.float32, torch.bfloat16):
        raise ValueError("dtype must be float32 or bfloat16")
    x = frames_u8.contiguous()
    f, h, w, _ = x.shape
    oh, ow = (int(size[0]), int(size[1])) if size is not None else (h, w)
    out = torch.empty((f, oh, ow, 3) if channels_last else (f, 3, oh, ow), device=x.device, dtype=dtype)
    m3, s3 = (C.c_float * 3)(*[float(v) for v in mean]), (C.c_float * 3)(*[float(v) for v in std])
    with torch.cuda.device(x.device):
        check(lib.lpb_frames_normalize(_ptr(x), f, h, w, oh, ow, m3, s3, int(bool(channels_last)), int(dtype == torch.bfloat16), _ptr(out), _stream()))
    return out


def pack_predictions(keypoints, confidences, table, cursor=None, row0: int = 0):
    """Write one chunk's (keypoints (T, 2K), confidences (T, K)) into rows of ``table`` (N, 3K) at the device cursor
    (int64 tensor, advanced by T) or at ``row0``.  Mutates ``table`` (and ``cursor``)."""
    kp = _cuda_f32(keypoints, "keypoints")
    cf = _cuda_f32(confidences, "confidences")
    t, k = cf.shape
    if table.dtype != torch.float32 or not table.is_contiguous() or table.shape[1] != 3 * k:
        raise ValueError(f"table must be contiguous fp32 (N, {3 * k}); got {tuple(table.shape)} {table.dtype}")
    with torch.cuda.device(kp.device):
        check(lib.lpb_pack_predictions(_ptr(kp), _ptr(cf), t, k, _ptr(table), table.shape[0], _ptr(cursor), int(row0), _stream()))
    return table


def plane_softmax_backward(probs: torch.Tensor, grad_probs: torch.Tensor) -> torch.Tensor:
    p = _cuda_f32(probs, "probs")
    g = _cuda_f32(grad_probs, "grad_probs")
    b, k, h, w = p.shape
    out = torch.empty_like(p)
    with torch.cuda.device(p.device):
        check(lib.lpb_plane_softmax_bwd(_ptr(p), _ptr(g), b * k, h * w, _ptr(out), _stream()))
    return out


# =====================================================================================
# heatmap losses
# =====================================================================================
class _HeatmapLoss(torch.autograd.Function):
    @staticmethod
    def forward(ctx, targets, preds, kind):
        b, k, h, w = preds.shape
        out = torch.empty((2,), device=preds.device, dtype=torch.float32)
        ws = torch.empty((b * k * 2,), device=preds.device, dtype=torch.float32)
        with torch.cuda.device(preds.device):
            check(lib.lpb_heatmap_loss_fwd(_ptr(targets), _ptr(preds), b * k, h, w, kind, _ptr(out), _ptr(ws), _stream()))
        ctx.save_for_backward(targets, preds, ws, out)
        ctx.kind = kind
        loss, count = out[0].clone(), out[1].clone()
        ctx.mark_non_differentiable(count)
        return loss, count

    @staticmethod
    def backward(ctx, g, _g_count):
        targets, preds, ws, out = ctx.saved_tensors
        b, k, h, w = preds.shape
        gp = torch.empty_like(preds)
        gg = g.reshape(1).contiguous().float()
        with torch.cuda.device(preds.device):
            check(lib.lpb_heatmap_loss_bwd(_ptr(targets), _ptr(preds), b * k, h, w, ctx.kind, _ptr(ws), _ptr(out), _ptr(gg), _ptr(gp), _stream()))
        return None, gp, None


def heatmap_loss(targets, preds, kind: str = "mse", return_count: bool = False):
    """mean heatmap loss over planes whose target is not all-zero; optionally also that plane count."""
    t = _cuda_f32(targets, "heatmaps_targ")
    p = _cuda_f32(preds, "heatmaps_pred")
    if t.shape != p.shape or t.dim() != 4:
        raise ValueError(f"heatmap shapes {tuple(t.shape)} vs {tuple(p.shape)}")
    loss, count = _HeatmapLoss.apply(t.detach(), p, _KIND[kind])
    return (loss, count) if return_count else loss


class _HeatmapMseFromKeypoints(torch.autograd.Function):
    @staticmethod
    def forward(ctx, kp, vis, preds, height, width, sigma):
        b, k, oh, ow = preds.shape
        out = torch.empty((2,), device=preds.device, dtype=torch.float32)
        ws = torch.empty((b * k * 2,), device=preds.device, dtype=torch.float32)
        with torch.cuda.device(preds.device):
            check(lib.lpb_heatmap_mse_from_keypoints_fwd(_ptr(kp), _ptr(vis), _ptr(preds), b * k, height, width, oh, ow, sigma, _ptr(out), _ptr(ws), _stream()))
        ctx.save_for_backward(kp, vis, preds, out)
        ctx.meta = (height, width, sigma)
        return out[0].clone()

    @staticmethod
    def backward(ctx, g):
        kp, vis, preds, out = ctx.saved_tensors
        height, width, sigma = ctx.meta
        b, k, oh, ow = preds.shape
        gp = torch.empty_like(preds)
        gg = g.reshape(1).contiguous().float()
        with torch.cuda.device(preds.device):
            check(lib.lpb_heatmap_mse_from_keypoints_bwd(_ptr(kp), _ptr(vis), _ptr(preds), b * k, height, width, oh, ow, sigma, _ptr(out), _ptr(gg), _ptr(gp), _stream()))
        return None, None, gp, None, None, None


def heatmap_mse_from_keypoints(keypoints, preds, height, width, sigma=1.25, visibility=None):
    """Fused target generation + HeatmapMSELoss (targets never written to HBM); differentiable in ``preds``."""
    kp = _cuda_f32(keypoints, "keypoints").detach()
    p = _cuda_f32(preds, "heatmaps_pred")
    vis = visibility.to(torch.int32).contiguous() if visibility is not None else None
    return _HeatmapMseFromKeypoints.apply(kp, vis, p, float(height), float(width), float(sigma))


class _TemporalHeatmapLoss(torch.autograd.Function):
    @staticmethod
    def forward(ctx, hm, cf, eps, kind, prob_threshold):
        t, k, h, w = hm.shape
        out = torch.empty((1,), device=hm.device, dtype=torch.float32)
        ws = torch.empty((max(t - 1, 1) * k,), device=hm.device, dtype=torch.float32)
        with torch.cuda.device(hm.device):
            check(lib.lpb_temporal_heatmap_loss_fwd(_ptr(hm), _ptr(cf), t, k, h, w, kind, _ptr(eps), float(prob_threshold), _ptr(out), _ptr(ws), _stream()))
        ctx.save_for_backward(hm, cf, eps, ws)
        ctx.meta = (kind, float(prob_threshold))
        return out[0]

    @staticmethod
    def backward(ctx, g):
        hm, cf, eps, ws = ctx.saved_tensors
        kind, thr = ctx.meta
        t, k, h, w = hm.shape
        gh = torch.empty_like(hm)
        gg = g.reshape(1).contiguous().float()
        with torch.cuda.device(hm.device):
            check(lib.lpb_temporal_heatmap_loss_bwd(_ptr(hm), _ptr(cf), _ptr(ws), t, k, h, w, kind, _ptr(eps), thr, _ptr(gg), _ptr(gh), _stream()))
        return gh, None, None, None, None


def temporal_heatmap_loss(heatmaps, confidences, kind: str, epsilon: torch.Tensor, prob_threshold: float):
    """TemporalHeatmapLoss (mse | kl) over consecutive frames; differentiable in ``heatmaps``."""
    hm = _cuda_f32(heatmaps, "heatmaps_pred")
    cf = _cuda_f32(confidences, "confidences").detach()
    k = hm.shape[1]
    eps = _cuda_f32(epsilon.to(hm.device).reshape(-1).expand(k) if epsilon.numel() in (1, k) else epsilon, "epsilon")
    if hm.shape[0] < 2:
        raise ValueError("temporal heatmap loss needs at least two frames")
    return _TemporalHeatmapLoss.apply(hm, cf, eps, _KIND[kind], float(prob_threshold))


# =====================================================================================
# unsupervised losses on (T, K, 2)
# =====================================================================================
class PcaParams:
    """Device-side parameters of one PCA loss (mirror of ``lpb_pca_desc``)."""

    def __init__(self, kp_index, n_sel, n_views, centering, mean, kept, epsilon, device):
        self.kp_index = torch.as_tensor(kp_index, dtype=torch.int32, device=device).contiguous()
        self.mean = torch.as_tensor(mean, dtype=torch.float32, device=device).contiguous()
        self.kept = torch.as_tensor(kept, dtype=torch.float32, device=device).contiguous().reshape(-1, self.mean.numel())
        self.n_sel, self.n_views = int(n_sel), int(n_views)
        self.centering = {None: 0, "mean": 1, "median": 2}[centering]
        self.epsilon = float(epsilon)
        d = 2 * self.n_views if self.n_views > 0 else 2 * self.n_sel
        if self.mean.numel() != d or self.kept.shape[1] != d:
            raise ValueError(f"PCA parameter dimension mismatch: D={d}, mean={self.mean.numel()}, kept={tuple(self.kept.shape)}")

    def desc(self) -> PcaDesc:
        return PcaDesc(self.kp_index.data_ptr(), self.n_sel, self.n_views, self.centering, self.kept.shape[0],
                       self.mean.data_ptr(), self.kept.data_ptr(), self.epsilon)


class _UnsupLosses(torch.autograd.Function):
    @staticmethod
    def forward(ctx, keypoints, confidences, temporal_eps, prob_threshold, temporal_on, sv, mv):
        n_clips, t, k2 = keypoints.shape
        out = torch.empty((n_clips, 4), device=keypoints.device, dtype=torch.float32)
        dsv = sv.desc() if sv is not None else None
        dmv = mv.desc() if mv is not None else None
        with torch.cuda.device(keypoints.device):
            check(lib.lpb_unsup_losses_fwd(_ptr(keypoints), _ptr(confidences), n_clips, t, k2 // 2, _ptr(temporal_eps), float(prob_threshold), int(temporal_on),
                                           C.byref(dsv) if dsv else None, C.byref(dmv) if dmv else None, _ptr(out), _stream()))
        ctx.save_for_backward(keypoints, confidences, temporal_eps)
        ctx.meta = (float(prob_threshold), int(temporal_on), sv, mv)
        return out

    @staticmethod
    def backward(ctx, g):
        keypoints, confidences, temporal_eps = ctx.saved_tensors
        thr, temporal_on, sv, mv = ctx.meta
        n_clips, t, k2 = keypoints.shape
        gk = torch.empty_like(keypoints)
        dsv = sv.desc() if sv is not None else None
        dmv = mv.desc() if mv is not None else None
        gg = g.contiguous().float()
        with torch.cuda.device(keypoints.device):
            check(lib.lpb_unsup_losses_bwd(_ptr(keypoints), _ptr(confidences), n_clips, t, k2 // 2, _ptr(temporal_eps), thr, temporal_on,
                                           C.byref(dsv) if dsv else None, C.byref(dmv) if dmv else None, _ptr(gg), _ptr(gk), _stream()))
        return gk, None, None, None, None, None, None


def unsup_losses(keypoints, confidences=None, temporal_eps=None, prob_threshold=0.0, pca_singleview: PcaParams | None = None,
                 pca_multiview: PcaParams | None = None):
    """One launch for the unsupervised loss stack.  keypoints (n_clips, T, 2K) or (T, 2K).

    Returns (n_clips, 4) = [temporal, pca_singleview, pca_multiview, 0] per clip (or (4,) for 2-D input).
    """
    kp = _cuda_f32(keypoints, "keypoints_pred")
    squeeze = kp.dim() == 2
    if squeeze:
        kp = kp[None]
    cf = None
    if confidences is not None:
        cf = _cuda_f32(confidences, "confidences").reshape(kp.shape[0], kp.shape[1], -1)
    k = kp.shape[2] // 2
    te = None
    if temporal_eps is not None:
        te = torch.as_tensor(temporal_eps, dtype=torch.float32, device=kp.device).reshape(-1)
        te = (te.expand(k) if te.numel() == 1 else te).contiguous()
        if te.numel() != k:
            raise ValueError(f"temporal epsilon has {te.numel()} entries for {k} keypoints")
    out = _UnsupLosses.apply(kp, cf, te, prob_threshold, te is not None, pca_singleview, pca_multiview)
    return out[0] if squeeze else out
