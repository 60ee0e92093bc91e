"""Heatmap head on B200: drop-in for ``lightning_pose.models.heads.heatmap``.

Same public surface as the reference module (``lightning_pose/models/heads/heatmap.py``):
``make_upsampling_layers`` (:20-71), ``initialize_upsampling_layers`` (:74-83), ``upsample``
(:86-100), ``run_subpixelmaxima`` (:103-144) and ``HeatmapHead`` (:147-227) with identical
constructor arguments, attributes and state-dict keys (``upsampling_layers.<i>.weight`` in
ConvTranspose2d ``(C_in, C_out, 3, 3)`` layout), so reference checkpoints load unchanged.

What differs is what runs: the ``nn.Sequential`` is only a parameter container.  ``forward``
hands its weights to the fused CUDA head (PixelShuffle folded into the first transposed
convolution's staging, gather-form deconvs, in-place plane softmax) and ``run_subpixelmaxima``
calls the fused soft-argmax decode, which never materialises the 4x-upsampled field.
"""
from __future__ import annotations

import math

import torch
from torch import nn

from lightning_pose_b200 import ops
from lightning_pose_b200.models.backbones import BACKBONE_STRIDES

__all__: list[str] = []


def make_upsampling_layers(in_channels: int, out_channels: int, int_channels: int, n_layers: int) -> nn.Sequential:
    """``PixelShuffle(2)`` followed by ``n_layers`` stride-2 3x3 transposed convolutions.

    Channel plan (reference :44-71): the shuffle divides channels by 4; intermediate layers use
    ``int_channels``; the last layer emits ``out_channels``.
    """
    widths = [in_channels // 4] + [int_channels] * (n_layers - 1) + [out_channels]
    layers: list[nn.Module] = [nn.PixelShuffle(2)]
    for c_in, c_out in zip(widths[:-1], widths[1:]):
        layers.append(
            nn.ConvTranspose2d(c_in, c_out, kernel_size=(3, 3), stride=(2, 2), padding=(1, 1), output_padding=(1, 1))
        )
    return nn.Sequential(*layers)


def initialize_upsampling_layers(layers: nn.Sequential) -> None:
    """xavier-uniform (gain 0.01) weights and zero bias for the deconvs (reference :74-83)."""
    for layer in list(layers)[1:]:
        if isinstance(layer, nn.ConvTranspose2d):
            nn.init.xavier_uniform_(layer.weight, gain=0.01)
            nn.init.zeros_(layer.bias)
        elif isinstance(layer, nn.BatchNorm2d):
            nn.init.ones_(layer.weight)
            nn.init.zeros_(layer.bias)


def upsample(inputs: torch.Tensor) -> torch.Tensor:
    """One 2x stage: bicubic (align_corners=False) + zero-padded 5x5 binomial blur (reference :86-100)."""
    return ops.upsample2x(inputs)


def run_subpixelmaxima(
    heatmaps: torch.Tensor, downsample_factor: int, temperature: torch.Tensor | float
) -> tuple[torch.Tensor, torch.Tensor]:
    """Soft-argmax decode: (batch, 2*num_keypoints) keypoints and (batch, num_keypoints) confidences."""
    return ops.decode_softargmax(heatmaps, int(downsample_factor), float(temperature))


class _HeadFunction(torch.autograd.Function):
    """Forward: fused CUDA head.  Backward: native tcgen05 kernels for the bf16 head (see ``HeatmapHead``)."""

    @staticmethod
    def forward(ctx, features, final_softmax, *params):
        n = len(params) // 2
        weights, biases = list(params[:n]), list(params[n:])
        saved = None
        out = None
        if features.dtype == torch.bfloat16 and n == 2 and features.is_cuda:
            res = ops._head_forward_bf16(features.contiguous(), weights, biases, final_softmax, train=True)
            if res is not None:
                out, saved = res
        if out is None:
            out = ops.head_forward(features, weights, biases, final_softmax)
        ctx.save_for_backward(features, out, *params)
        ctx.final_softmax, ctx.n, ctx.saved = final_softmax, n, saved
        return out

    @staticmethod
    def backward(ctx, g):
        features, out, *params = ctx.saved_tensors
        n = ctx.n
        weights, biases = params[:n], params[n:]
        g = g.contiguous().float()
        if ctx.final_softmax:  # d softmax: p * (g - sum(g * p)) per plane
            g = ops.plane_softmax_backward(out, g)
        if ctx.saved is not None:
            dfeat, dw1, db1, dw2, db2 = ops.head_backward_bf16(
                g, ctx.saved, tuple(features.shape), weights[0], weights[1], need_dfeat=ctx.needs_input_grad[0]
            )
            ctx.saved = None
            return (dfeat, None, dw1.to(weights[0].dtype), dw2.to(weights[1].dtype), db1.to(biases[0].dtype), db2.to(biases[1].dtype))
        # shapes outside the tensor-core tiling and the fp32 head: transposed-conv dgrad/wgrad through the
        # framework's conv ops ("backward")
        cdt = torch.bfloat16 if features.dtype == torch.bfloat16 else torch.float32  # same precision as the forward
        with torch.enable_grad():
            f = features.detach().to(cdt).requires_grad_(ctx.needs_input_grad[0])
            ps = [p.detach().float().requires_grad_(True) for p in params]
            x = torch.nn.functional.pixel_shuffle(f, 2)
            for wt, bs in zip(ps[:n], ps[n:]):
                x = torch.nn.functional.conv_transpose2d(x, wt.to(cdt), bs.to(cdt), stride=2, padding=1, output_padding=1)
            ins = ([f] if ctx.needs_input_grad[0] else []) + ps
            grads = torch.autograd.grad(x, ins, g.to(cdt))
        gf = grads[0] if ctx.needs_input_grad[0] else None
        gp = grads[1:] if ctx.needs_input_grad[0] else grads
        return (gf, None, *gp)


class HeatmapHead(nn.Module):
    """Deconvolution head: backbone features -> per-keypoint spatial-softmax heatmaps.

    Constructor / attributes follow the reference (:155-201): ``n_layers = log2(stride) -
    downsample_factor - 1`` deconvs after the PixelShuffle, soft-argmax temperature 1000.
    """

    def __init__(
        self,
        backbone_arch: str,
        in_channels: int,
        out_channels: int,
        deconv_out_channels: int | None = None,
        downsample_factor: int = 2,
        final_softmax: bool = True,
    ) -> None:
        super().__init__()
        self.backbone_arch = backbone_arch
        self.in_channels = in_channels
        self.out_channels = out_channels
        self.deconv_out_channels = deconv_out_channels
        self.downsample_factor = downsample_factor
        self.final_softmax = final_softmax
        self.temperature = torch.tensor(1000.0)
        stride = BACKBONE_STRIDES.get(backbone_arch, 32)
        n_layers = int(math.log2(stride)) - downsample_factor - 1
        self.upsampling_layers = make_upsampling_layers(
            in_channels=in_channels,
            out_channels=out_channels,
            int_channels=deconv_out_channels or out_channels,
            n_layers=n_layers,
        )
        initialize_upsampling_layers(self.upsampling_layers)

    def _deconvs(self) -> list[nn.ConvTranspose2d]:
        return [m for m in self.upsampling_layers if isinstance(m, nn.ConvTranspose2d)]

    def forward(self, features: torch.Tensor) -> torch.Tensor:
        deconvs = self._deconvs()
        params = [d.weight for d in deconvs] + [d.bias for d in deconvs]
        return _HeadFunction.apply(features, bool(self.final_softmax), *params)

    def run_subpixelmaxima(self, heatmaps: torch.Tensor) -> tuple[torch.Tensor, torch.Tensor]:
        return run_subpixelmaxima(heatmaps, self.downsample_factor, self.temperature)
