"""SoundStream VQ-VAE -- host-side mirror of the reference ``models/vq_brain.py``.

Same class names, constructor kwargs, ``forward`` signatures, return conventions and state-dict
keys as the reference (models/vq_brain.py:22-243), so this module drops into
``notebooks_trainer/vq_brain_trainer.ipynb`` and ``utils/train_utils.py:138``.  What changes
underneath:

* the quantiser is ``frankenstein_b200.vector_quantize.VectorQuantize`` (tcgen05 search + fused
  epilogue kernels) instead of the third-party ``vector_quantize_pytorch`` module;
* ``custom_l1_loss`` is one fused masked-sum kernel (no ``nonzero`` host sync, no dynamic shape);
* ``calculate_perp`` is computed from the code histogram (no ``[N, K]`` one-hot).

* the causal conv / transposed-conv stacks run channels-last on the library's tcgen05
  GEMMs as implicit GEMMs (``conv.py``: ``[B, T, C]`` activations, im2col rows are overlapping views of the padded
  signal) whenever every layer's channel counts are multiples of 64; other widths, and ``FK_CONV=cudnn``, use cuDNN's
  NHWC kernels on a 4-D view (no permute copies either way).
"""
from __future__ import annotations

import os

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import _lib, conv, gemm
from ._lib import DTYPE_CODE, FkError, check, counters, lib, on_tensor_device, ptr, require_cuda, require_device, stream
from .vector_quantize import VectorQuantize


# "own" (default): the conv stacks run channels-last on the library's tcgen05 GEMMs (conv.py: im2col rows are overlapping
# views of the padded signal); "cudnn": the same stacks as 1 x k conv2d on cuDNN's NHWC kernels (A/B comparison only).
CONV_IMPL = os.environ.get("FK_CONV", "own")
_CL3D = [False]      # set by _ChannelsFirstStack while its layers run on [B, T, C] activations


def _own_conv_ok(m) -> bool:
    return (m.padding_mode == "zeros" and conv.conv_supported(m.in_channels, m.out_channels, m.kernel_size[0], m.stride[0],
                                                             m.dilation[0], m.groups))


class CausalConv1d(nn.Conv1d):
    """models/vq_brain.py:22-28 -- left padding of dilation * (k - 1)."""

    def __init__(self, *args, **kwargs):
        super().__init__(*args, **kwargs)
        self.causal_padding = self.dilation[0] * (self.kernel_size[0] - 1)

    def forward(self, x):
        if _CL3D[0] and x.dim() == 3:                       # [B, T, C]: implicit GEMM on tcgen05 (conv.py)
            return conv.causal_conv1d_cl(x, self.weight, self.bias, self.stride[0])
        if x.dim() == 4:
            # channels-last path ([B, C, 1, T] with NHWC strides, see _ChannelsFirstStack): the same convolution as a
            # 1 x k conv2d, so that cuDNN runs its NHWC kernels without a layout conversion before and after
            return F.conv2d(F.pad(x, [self.causal_padding, 0]), self.weight.unsqueeze(2), self.bias,
                            stride=(1, self.stride[0]), dilation=(1, self.dilation[0]), groups=self.groups)
        return self._conv_forward(F.pad(x, [self.causal_padding, 0]), self.weight, self.bias)


class CausalConvTranspose1d(nn.ConvTranspose1d):
    """models/vq_brain.py:31-45 -- transposed conv with the trailing k - s samples trimmed."""

    def __init__(self, *args, **kwargs):
        super().__init__(*args, **kwargs)
        self.causal_padding = (self.dilation[0] * (self.kernel_size[0] - 1) + self.output_padding[0] + 1
                               - self.stride[0])

    def forward(self, x, output_size=None):
        if self.padding_mode != 'zeros':
            raise ValueError('Only `zeros` padding mode is supported for ConvTranspose1d')
        if _CL3D[0] and x.dim() == 3:
            return conv.causal_conv_transpose1d_cl(x, self.weight, self.bias)
        if x.dim() == 4:                                    # channels-last path (see CausalConv1d)
            y = F.conv_transpose2d(x, self.weight.unsqueeze(2), self.bias, stride=(1, self.stride[0]),
                                   padding=(0, self.padding[0]), output_padding=(0, self.output_padding[0]),
                                   groups=self.groups, dilation=(1, self.dilation[0]))
            return y[..., :-self.causal_padding]
        output_padding = self._output_padding(x, output_size, self.stride, self.padding, self.kernel_size,
                                              self.dilation)
        y = F.conv_transpose1d(x, self.weight, self.bias, self.stride, self.padding, output_padding, self.groups,
                               self.dilation)
        return y[..., :-self.causal_padding]


class _PointwiseConv1d(nn.Conv1d):
    """nn.Conv1d (same parameters / state-dict keys) that also takes the channels-last 4-D layout."""

    def forward(self, x):
        if _CL3D[0] and x.dim() == 3:                       # 1 x 1 convolution = the plain projection of the channel axis
            return gemm.linear(x, self.weight.squeeze(2), self.bias)
        if x.dim() == 4:
            return F.conv2d(x, self.weight.unsqueeze(2), self.bias, stride=(1, self.stride[0]),
                            padding=(0, self.padding[0]), dilation=(1, self.dilation[0]), groups=self.groups)
        return super().forward(x)


def _interleave_elu(mods):
    """[m0, ELU, m1, ELU, ..., mN] -- the Sequential indexing the reference checkpoints expect."""
    out = []
    for i, m in enumerate(mods):
        if i:
            out.append(nn.ELU())
        out.append(m)
    return nn.Sequential(*out)


class ResidualUnit(nn.Module):
    """x + Conv1x1(ELU(CausalConv3(x)))  (models/vq_brain.py:48-63)."""

    def __init__(self, in_channels, out_channels, dilation):
        super().__init__()
        self.dilation = dilation
        self.layers = _interleave_elu([
            CausalConv1d(in_channels, out_channels, kernel_size=3, dilation=dilation),
            _PointwiseConv1d(out_channels, in_channels, kernel_size=1),
        ])

    def forward(self, x):
        return x + self.layers(x)


def _residual_stack(ch, n=3):
    return [ResidualUnit(ch, ch, dilation=1) for _ in range(n)]


class EncoderBlock(nn.Module):
    """3 residual units then a strided causal conv (k = 2*stride)  (models/vq_brain.py:66-90)."""

    def __init__(self, in_channels, out_channels, stride):
        super().__init__()
        self.layers = _interleave_elu(_residual_stack(in_channels) + [
            CausalConv1d(in_channels, out_channels, kernel_size=2 * stride, stride=stride)])

    def forward(self, x):
        return self.layers(x)


class DecoderBlock(nn.Module):
    """strided causal transposed conv then 3 residual units  (models/vq_brain.py:93-117)."""

    def __init__(self, in_channels, out_channels, stride):
        super().__init__()
        self.layers = _interleave_elu([
            CausalConvTranspose1d(in_channels, out_channels, kernel_size=2 * stride, stride=stride)]
            + _residual_stack(out_channels))

    def forward(self, x):
        return self.layers(x)


class _ChannelsFirstStack(nn.Module):
    """[B, T, C] in and out.  The reference permutes to [B, C, T] (vq_brain.py:135-137, 156-158) and cuDNN then converts
    to NHWC and back around every convolution.  Here [B, T, C] is viewed, without a copy, as the channels-last 4-D tensor
    [B, C, 1, T] (its NHWC strides are exactly those of a contiguous [B, T, C]), every layer is the equivalent 1 x k
    conv2d, and the result is viewed back: no permute copies and no layout-conversion kernels."""

    def _own_ok(self) -> bool:
        if getattr(self, "_own_checked", None) is None:
            ok = True
            for m in self.modules():
                if isinstance(m, CausalConv1d):
                    ok = ok and _own_conv_ok(m)
                elif isinstance(m, CausalConvTranspose1d):
                    ok = ok and (m.kernel_size[0] == 4 and m.stride[0] == 2 and m.dilation[0] == 1 and m.groups == 1
                                 and m.output_padding[0] == 0 and m.in_channels % 32 == 0 and m.out_channels % 32 == 0)
                elif isinstance(m, _PointwiseConv1d):
                    ok = ok and (m.kernel_size[0] == 1 and m.stride[0] == 1 and m.groups == 1
                                 and gemm.linear_supported(m.in_channels, m.out_channels))
            self._own_checked = ok
        return self._own_checked

    def forward(self, x):
        if x.is_cuda and x.dim() == 3 and CONV_IMPL == "own" and self._own_ok():
            # activations stay [B, T, C] bf16: every convolution is an implicit GEMM over overlapping-row views (conv.py),
            # no permutes, no im2col copies, no cuDNN
            _CL3D[0] = True
            try:
                return self.layers(x)
            finally:
                _CL3D[0] = False
        if x.is_cuda and x.dim() == 3 and x.is_contiguous():
            y = self.layers(x.transpose(1, 2).unsqueeze(2))          # [B, C, 1, T], channels_last strides
            return y.squeeze(2).transpose(1, 2)
        return self.layers(x.transpose(1, 2)).transpose(1, 2)


class Encoder(_ChannelsFirstStack):
    """models/vq_brain.py:120-138: conv k5 -> 2 x EncoderBlock(stride 2) -> conv k3; T/4 tokens of width D."""

    def __init__(self, C, D, n_electrodes):
        super().__init__()
        self.layers = _interleave_elu([
            CausalConv1d(n_electrodes, C, kernel_size=5),
            EncoderBlock(C, C, stride=2),
            EncoderBlock(C, C, stride=2),
            CausalConv1d(C, D, kernel_size=3),
        ])


class Decoder(_ChannelsFirstStack):
    """models/vq_brain.py:141-159: mirror of the encoder."""

    def __init__(self, C, D, n_channels_out):
        super().__init__()
        self.layers = _interleave_elu([
            CausalConv1d(D, C, kernel_size=3),
            DecoderBlock(C, C, stride=2),
            DecoderBlock(C, C, stride=2),
            CausalConv1d(C, n_channels_out, kernel_size=5),
        ])


class _MaskedL1(torch.autograd.Function):
    """custom_l1_loss (models/vq_brain.py:220-227) as one fused forward and one fused backward kernel."""

    @staticmethod
    @on_tensor_device
    def forward(ctx, pred, gt):
        require_cuda(pred, gt)
        require_device()
        ctx.pred_dtype = pred.dtype
        if pred.dtype not in (torch.float32, torch.bfloat16):
            pred = pred.float()
        B, T, C = pred.shape
        p = pred.contiguous()
        g = gt.contiguous().float()
        R = B * T
        nb = lib().fk_masked_l1_partials(R)
        dev = pred.device
        row_valid = torch.empty(R, device=dev, dtype=torch.uint8)
        ps = torch.empty(nb, device=dev, dtype=torch.float32)
        pc = torch.empty(nb, device=dev, dtype=torch.float32)
        loss = torch.empty(1, device=dev, dtype=torch.float32)
        denom = torch.empty(1, device=dev, dtype=torch.float32)
        check(lib().fk_masked_l1_forward(ptr(p), DTYPE_CODE[p.dtype], ptr(g), R, C, ptr(row_valid), ptr(ps), ptr(pc),
                                         counters(_lib.CTR_MASKED_L1), ptr(loss), ptr(denom), stream()), "fk_masked_l1_forward")
        ctx.save_for_backward(p, g, row_valid, denom)
        return loss.view(())

    @staticmethod
    @on_tensor_device
    def backward(ctx, g_loss):
        p, g, row_valid, denom = ctx.saved_tensors
        B, T, C = p.shape
        grad = torch.empty_like(p)
        gl = g_loss.reshape(1).contiguous().float()
        check(lib().fk_masked_l1_backward(ptr(p), DTYPE_CODE[p.dtype], ptr(g), ptr(row_valid), ptr(gl), ptr(denom), B * T,
                                          C, ptr(grad), stream()), "fk_masked_l1_backward")
        return grad.to(ctx.pred_dtype), None


@on_tensor_device
def perplexity(indices: torch.Tensor, codebook_size: int) -> torch.Tensor:
    """calculate_perp (models/vq_brain.py:238-243) from the code histogram."""
    require_cuda(indices)
    require_device()
    ind = indices.reshape(-1).contiguous()
    bins = torch.empty(codebook_size, device=ind.device, dtype=torch.float32)
    out = torch.empty(1, device=ind.device, dtype=torch.float32)
    check(lib().fk_vq_perplexity(ptr(ind), ind.numel(), codebook_size, ptr(bins), ptr(out), stream()), "fk_vq_perplexity")
    return out.view(())


class SoundStream(nn.Module):
    """models/vq_brain.py:162-243.  forward(x[B,T,n_electrodes]) -> (rec_loss + commit_loss [1], o [B,T,n_electrodes])."""

    def __init__(self, C, D, codebook_size, n_electrodes, use_cosine_sim=True):
        super().__init__()
        self.codebook_size = codebook_size
        self.encoder = Encoder(C=C, D=D, n_electrodes=n_electrodes)
        self.quantizer = VectorQuantize(
            dim=D,
            codebook_size=codebook_size,
            commitment_weight=0.25,
            channel_last=True,
            kmeans_init=True,
            threshold_ema_dead_code=2,
            use_cosine_sim=use_cosine_sim,
        )
        self.decoder = Decoder(C=C, D=D, n_channels_out=n_electrodes)
        self.compute_perplexity = False     # the reference computes it and discards the value (vq_brain.py:212)
        self.last_perplexity = None

    def forward(self, x, targets=None, date_info=None):
        e = self.encoder(x)
        quantized, indices, commit_loss = self.quantizer(e)
        if quantized.dtype != e.dtype and not torch.is_autocast_enabled():
            quantized = quantized.to(e.dtype)      # module cast to bf16/fp16 without autocast
        o = self.decoder(quantized)
        if self.compute_perplexity:
            self.last_perplexity = self.calculate_perp(indices)
        rec_loss = self.custom_l1_loss(o, x)
        total_loss = rec_loss + commit_loss
        return total_loss, o

    def custom_l1_loss(self, pred, gt):
        return _MaskedL1.apply(pred, gt)

    def get_quantize_vectors(self, x):
        e = self.encoder(x)
        quantized, indices, commit_loss = self.quantizer(e)
        return indices, quantized

    def calculate_perp(self, indices):
        return perplexity(indices, self.codebook_size)
