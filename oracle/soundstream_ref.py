"""CPU fp32 restatement of the reference SoundStream VQ-VAE -- TEST INFRASTRUCTURE ONLY.

Restates ``models/vq_brain.py`` of the reference as plain functions over a state_dict (the
reference's own parameter names), so that it runs on the GPU box where ``/root/reference`` does not
exist.  Pinned against the real reference in the build container: ``scripts/make_golden.py`` runs the
unmodified ``models/vq_brain.SoundStream`` (through ``oracle/ref_shims.py``) and this file on the
same weights and inputs and stores the reference outputs under ``tests/golden/``;
``tests/test_oracle_cpu.py`` re-checks this file against those fixtures on every run.

The quantiser inside is ``oracle.vector_quantize_ref.VectorQuantizeRef`` (parity unpinned -- see
that file's header: the library is an un-vendored dependency of the reference).

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline legs import this.
"""
from __future__ import annotations

import torch
import torch.nn.functional as F

from oracle.vector_quantize_ref import VectorQuantizeRef


def _causal_conv(x, sd, prefix, dilation=1, stride=1):
    """CausalConv1d.forward (models/vq_brain.py:22-28): left pad dilation*(k-1), then conv."""
    w, b = sd[prefix + ".weight"], sd[prefix + ".bias"]
    pad = dilation * (w.shape[-1] - 1)
    return F.conv1d(F.pad(x, [pad, 0]), w, b, stride=stride, dilation=dilation)


def _causal_conv_transpose(x, sd, prefix, stride):
    """CausalConvTranspose1d.forward (models/vq_brain.py:31-45): trim k - stride trailing samples."""
    w, b = sd[prefix + ".weight"], sd[prefix + ".bias"]
    trim = w.shape[-1] - stride
    return F.conv_transpose1d(x, w, b, stride=stride)[..., :-trim]


def _residual_unit(x, sd, prefix):
    """ResidualUnit (models/vq_brain.py:48-63): x + conv1x1(elu(causal_conv3(x)))."""
    y = _causal_conv(x, sd, prefix + ".layers.0")
    y = F.conv1d(F.elu(y), sd[prefix + ".layers.2.weight"], sd[prefix + ".layers.2.bias"])
    return x + y


def _encoder_block(x, sd, prefix, stride=2):
    """EncoderBlock (models/vq_brain.py:66-90)."""
    for i in (0, 2, 4):
        x = F.elu(_residual_unit(x, sd, f"{prefix}.layers.{i}"))
    return _causal_conv(x, sd, f"{prefix}.layers.6", stride=stride)


def _decoder_block(x, sd, prefix, stride=2):
    """DecoderBlock (models/vq_brain.py:93-117)."""
    x = F.elu(_causal_conv_transpose(x, sd, f"{prefix}.layers.0", stride))
    x = F.elu(_residual_unit(x, sd, f"{prefix}.layers.2"))
    x = F.elu(_residual_unit(x, sd, f"{prefix}.layers.4"))
    return _residual_unit(x, sd, f"{prefix}.layers.6")


def encoder_forward(sd, x):
    """Encoder.forward (models/vq_brain.py:120-138): [B,T,C_in] -> [B,T/4,D]."""
    h = x.transpose(1, 2)
    h = F.elu(_causal_conv(h, sd, "encoder.layers.0"))
    h = F.elu(_encoder_block(h, sd, "encoder.layers.2"))
    h = F.elu(_encoder_block(h, sd, "encoder.layers.4"))
    h = _causal_conv(h, sd, "encoder.layers.6")
    return h.transpose(1, 2)


def decoder_forward(sd, q):
    """Decoder.forward (models/vq_brain.py:141-159): [B,T/4,D] -> [B,T,C_out]."""
    h = q.transpose(1, 2)
    h = F.elu(_causal_conv(h, sd, "decoder.layers.0"))
    h = F.elu(_decoder_block(h, sd, "decoder.layers.2"))
    h = F.elu(_decoder_block(h, sd, "decoder.layers.4"))
    h = _causal_conv(h, sd, "decoder.layers.6")
    return h.transpose(1, 2)


def custom_l1_loss(pred, gt):
    """SoundStream.custom_l1_loss (models/vq_brain.py:220-227): mean |pred-gt| over rows with any non-zero gt."""
    real = ~torch.all(gt == 0, dim=2)
    return F.l1_loss(pred, gt, reduction="none")[real.nonzero(as_tuple=True)].mean()


def calculate_perp(indices, codebook_size):
    """SoundStream.calculate_perp (models/vq_brain.py:238-243)."""
    enc = F.one_hot(indices, codebook_size).float().reshape(-1, codebook_size)
    p = enc.mean(0)
    return (-(p * torch.log(p + 1e-10)).sum()).exp()


class SoundStreamRef:
    """SoundStream.forward (models/vq_brain.py:198-218) over detached parameter tensors that require grad."""

    def __init__(self, state_dict, D, codebook_size, use_cosine_sim, training=True, **vq_kwargs):
        self.sd = {k: v.detach().clone().float().requires_grad_(v.is_floating_point() and "quantizer" not in k)
                   for k, v in state_dict.items()}
        self.vq = VectorQuantizeRef(dim=D, codebook_size=codebook_size, commitment_weight=0.25, kmeans_init=True,
                                    threshold_ema_dead_code=2, use_cosine_sim=use_cosine_sim, **vq_kwargs)
        self.vq.load_state_dict({k[len("quantizer."):]: v for k, v in state_dict.items() if k.startswith("quantizer.")})
        self.vq.train(training)
        self.codebook_size = codebook_size

    def parameters(self):
        return [v for k, v in self.sd.items() if v.requires_grad]

    def forward(self, x):
        e = encoder_forward(self.sd, x)
        quantized, indices, commit = self.vq(e)
        o = decoder_forward(self.sd, quantized)
        self.last_indices = indices
        self.last_e = e
        return custom_l1_loss(o, x) + commit, o

    __call__ = forward
