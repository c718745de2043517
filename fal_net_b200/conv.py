"""Convolution front-end of the encoder-decoder and the VGG perceptual slices.

Activations are bf16 tensors of logical shape [B,C,H,W] in ``torch.channels_last`` memory format,
i.e. NHWC in memory -- the layout the tcgen05 implicit-GEMM kernels (csrc/conv_tc.cu) consume.

STATUS (round 1): the hand-written tcgen05 kernels are being brought up layer family by layer family;
every call that is not yet served natively goes through ``torch.nn.functional.conv2d`` (cuDNN, bf16,
fp32 accumulate) and is COUNTED in ``LIBRARY_CALLS`` so that bench.py can report which share of the
step still runs on library kernels.  This is a temporary bring-up path on the GPU, not a CPU fallback:
nothing here runs without CUDA.
"""
from __future__ import annotations

import torch
import torch.nn.functional as F

LIBRARY_CALLS = {"conv2d": 0}
CL = torch.channels_last


def input_to_nhwc(x: torch.Tensor) -> torch.Tensor:
    """fp32 NCHW image -> bf16 NHWC activation."""
    return x.to(dtype=torch.bfloat16, memory_format=CL)


def _act(y, act):
    if act == "elu":
        return F.elu(y)
    if act == "relu":
        return F.relu(y)
    return y


def conv3x3(x, weight, bias=None, stride=1, act=None, residual=None, const_channel=None, upsample_to=None, concat=None):
    """y = act(conv3x3(gather(x)) + bias [+ residual]) with pad 1.

    gather(x) = nearest-upsample to ``upsample_to`` (reference deconv, models/FAL_netB.py:58), then channel
    concat with ``concat`` (skip connection, :153-173) and/or with a per-sample constant plane
    ``const_channel`` [B] (the max_disp/100 "flow" channel, :145,208-209)."""
    if upsample_to is not None and tuple(x.shape[2:]) != tuple(upsample_to):
        x = F.interpolate(x, size=upsample_to, mode="nearest")
    if concat is not None:
        x = torch.cat((x, concat), 1)
    if const_channel is not None:
        B, _, H, W = x.shape
        plane = const_channel.to(x.dtype).view(B, 1, 1, 1).expand(B, 1, H, W)
        x = torch.cat((x, plane), 1)
    x = x.contiguous(memory_format=CL)
    LIBRARY_CALLS["conv2d"] += 1
    y = F.conv2d(x, weight.to(torch.bfloat16), None if bias is None else bias.to(torch.bfloat16), stride, 1)
    if residual is not None:
        y = y + residual
    return _act(y, act)


def fold_logit_conv(w_iconv1, w0, b0):
    """iconv1 (3x3, no bias, no activation; reference :127,174) followed by conv0 (1x1 + bias; :190,215)
    == one 3x3 conv with W'[o,c,kh,kw] = sum_m W0[o,m] * W_iconv1[m,c,kh,kw] and bias b0."""
    return torch.einsum("om,mckl->ockl", w0[:, :, 0, 0], w_iconv1), b0


def conv3x3_logits(u, skip, w_iconv1, w0, b0):
    """Last layer: concat(u, skip) -> folded 3x3 conv -> fp32 planar logits [B,N,H,W]."""
    w, b = fold_logit_conv(w_iconv1, w0, b0)
    x = torch.cat((u, skip), 1).contiguous(memory_format=CL)
    LIBRARY_CALLS["conv2d"] += 1
    y = F.conv2d(x, w.to(torch.bfloat16), None, 1, 1)
    return (y.float() + b.view(1, -1, 1, 1)).contiguous()


# ------------------------------------------------------------------------------------------------
# VGG19 features[0:19] (the three pooled activations of /root/reference/loss_functions.py:21-29,36-44)
# ------------------------------------------------------------------------------------------------
VGG_CFG = (64, 64, "M", 128, 128, "M", 256, 256, 256, 256, "M")


def vgg_features(ws, x):
    """ws: list of (weight fp32 [Co,Ci,3,3], bias fp32 [Co]) for the 8 convs; x bf16 NHWC (3 channels)."""
    outs, i = [], 0
    for v in VGG_CFG:
        if v == "M":
            x = F.max_pool2d(x, 2, 2)
            outs.append(x)
        else:
            w, b = ws[i]
            x = conv3x3(x, w, b, act="relu")
            i += 1
    return outs
