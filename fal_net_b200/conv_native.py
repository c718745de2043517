"""Thin torch wrappers over the tcgen05 implicit-GEMM convolution kernels (csrc/conv_tc.cu, csrc/conv_aux.cu).

Activations: bf16 tensors of logical shape [B,C,H,W] in channels_last memory format (= NHWC in memory)."""
from __future__ import annotations

import torch

from . import _lib

CL = torch.channels_last
ACT = {None: 0, "none": 0, "elu": 1, "relu": 2}


def pack_weight(w: torch.Tensor, cout_pad: int | None = None) -> torch.Tensor:
    """[Cout,Cin,3,3] (any float dtype) -> bf16 KRSC [Cout_pad,3,3,Cin], zero rows beyond Cout."""
    Cout, Cin = w.shape[0], w.shape[1]
    cout_pad = cout_pad or (Cout + 31) // 32 * 32
    out = torch.zeros(cout_pad, 3, 3, Cin, device=w.device, dtype=torch.bfloat16)
    out[:Cout] = w.detach().permute(0, 2, 3, 1).to(torch.bfloat16)
    return out


_BORDER_MASKS: dict = {}


def _border_masks(device) -> torch.Tensor:
    """[16,3,3] 0/1 masks: which taps of a 3x3 window stay inside the image for border class rc*4+cc
    (bit0: first tap outside, bit1: last tap outside)."""
    m = _BORDER_MASKS.get(str(device))
    if m is None:
        m = torch.zeros(16, 3, 3)
        for rc in range(4):
            for cc in range(4):
                for kh in range(3):
                    for kw in range(3):
                        out = (kh == 0 and rc & 1) or (kh == 2 and rc & 2) or (kw == 0 and cc & 1) or (kw == 2 and cc & 2)
                        m[rc * 4 + cc, kh, kw] = 0.0 if out else 1.0
        m = m.to(device)
        _BORDER_MASKS[str(device)] = m
    return m


def const_channel_table(wf: torch.Tensor, cout_pad=None) -> torch.Tensor:
    """wf [Cout,3,3]: weights of a spatially constant input channel -> [16,Cout] sums over the taps that stay inside
    the image, per border class (one small matmul on the device; no host-side indexing, so it is graph-capturable)."""
    return (_border_masks(wf.device).view(16, 9) @ wf.float().reshape(wf.shape[0], 9).t()).contiguous()


def _nhwc(t: torch.Tensor) -> torch.Tensor:
    assert t.dtype == torch.bfloat16 and t.is_cuda
    return t.contiguous(memory_format=CL)


def conv3x3_fwd(x, w_krsc, bias=None, stride=1, act=0, residual=None, x2=None, cout=None, planar_out=None, ctab=None,
                cscale=None):
    """x, x2, residual: bf16 [B,C,H,W] channels_last.  Returns bf16 [B,Cout,Ho,Wo] channels_last, or fills and returns
    ``planar_out`` (fp32 [B,Cout,Ho,Wo] with uniform row pitch) when given."""
    x = _nhwc(x)
    B, C1, H, W = x.shape
    C2 = 0
    if x2 is not None:
        x2 = _nhwc(x2)
        C2 = x2.shape[1]
        assert x2.shape[0] == B and x2.shape[2:] == x.shape[2:]
    cout_pad = w_krsc.shape[0]
    cout = cout or cout_pad
    assert w_krsc.shape[3] == C1 + C2 and w_krsc.dtype == torch.bfloat16 and w_krsc.is_contiguous()
    Ho, Wo = (H - 1) // stride + 1, (W - 1) // stride + 1
    if residual is not None:
        residual = _nhwc(residual)
        assert residual.shape == (B, cout, Ho, Wo)
    if bias is not None:
        bias = bias.detach().float().contiguous()
    if planar_out is not None:
        y, planar, pitch, out_c = planar_out, 1, planar_out.stride(2), cout
        assert planar_out.shape == (B, cout, Ho, Wo) and planar_out.dtype == torch.float32
    else:
        y = torch.empty((B, cout, Ho, Wo), device=x.device, dtype=torch.bfloat16, memory_format=CL)
        planar, pitch, out_c = 0, 0, cout
    rc = _lib.lib().faln_conv3x3_fwd(_lib.ptr(x), _lib.ptr(x2), _lib.ptr(w_krsc), _lib.ptr(bias), _lib.ptr(ctab),
                                     _lib.ptr(cscale), _lib.ptr(residual), _lib.ptr(y), B, H, W, C1, C2, cout, cout_pad,
                                     stride, int(act), planar, pitch, out_c, _lib.cur_stream())
    _lib.check(rc, "faln_conv3x3_fwd")
    return y


def stem_conv(x, w, bias, act, flip_x=False):
    """fp32 NCHW image [B,3,H,W] -> bf16 channels_last [B,Cout,H,W]; w [Cout,3,3,3] fp32."""
    x = _lib.f32c(x)
    B, _, H, W = x.shape
    Cout = w.shape[0]
    y = torch.empty((B, Cout, H, W), device=x.device, dtype=torch.bfloat16, memory_format=CL)
    rc = _lib.lib().faln_stem_conv(_lib.ptr(x), _lib.ptr(w.detach().float().contiguous()),
                                   _lib.ptr(None if bias is None else bias.detach().float().contiguous()), _lib.ptr(y), B, H,
                                   W, Cout, int(act), int(flip_x), _lib.cur_stream())
    _lib.check(rc, "faln_stem_conv")
    return y


def upsample_nearest(x, size):
    x = _nhwc(x)
    B, C, Hi, Wi = x.shape
    Ho, Wo = size
    if (Hi, Wi) == (Ho, Wo):
        return x
    y = torch.empty((B, C, Ho, Wo), device=x.device, dtype=torch.bfloat16, memory_format=CL)
    rc = _lib.lib().faln_upsample_nearest_nhwc(_lib.ptr(x), _lib.ptr(y), B, Hi, Wi, Ho, Wo, C, _lib.cur_stream())
    _lib.check(rc, "faln_upsample_nearest_nhwc")
    return y


def maxpool2(x):
    x = _nhwc(x)
    B, C, Hi, Wi = x.shape
    y = torch.empty((B, C, Hi // 2, Wi // 2), device=x.device, dtype=torch.bfloat16, memory_format=CL)
    rc = _lib.lib().faln_maxpool2_nhwc(_lib.ptr(x), _lib.ptr(y), B, Hi, Wi, C, _lib.cur_stream())
    _lib.check(rc, "faln_maxpool2_nhwc")
    return y
