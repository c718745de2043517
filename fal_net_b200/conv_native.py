"""Thin torch wrappers over the tcgen05 implicit-GEMM convolution kernels (csrc/conv_tc.cu)."""
from __future__ import annotations

import torch

from . import _lib

CL = torch.channels_last


def pack_weight(w: torch.Tensor, cout_pad: int | None = None) -> torch.Tensor:
    """[Cout,Cin,3,3] (any float dtype) -> bf16 KRSC [Cout_pad,3,3,Cin], zero rows beyond Cout."""
    Cout, Cin = w.shape[0], w.shape[1]
    cout_pad = cout_pad or (Cout + 31) // 32 * 32
    out = torch.zeros(cout_pad, 3, 3, Cin, device=w.device, dtype=torch.bfloat16)
    out[:Cout] = w.detach().permute(0, 2, 3, 1).to(torch.bfloat16)
    return out


def _nhwc(t: torch.Tensor) -> torch.Tensor:
    assert t.dtype == torch.bfloat16 and t.is_cuda
    return t.contiguous(memory_format=CL)


def conv3x3_fwd(x, w_krsc, bias=None, stride=1, act=0, residual=None, x2=None, cout=None, planar_out=None):
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
    rc = _lib.lib().faln_conv3x3_fwd(_lib.ptr(x), _lib.ptr(x2), _lib.ptr(w_krsc), _lib.ptr(bias), _lib.ptr(residual),
                                     _lib.ptr(y), B, H, W, C1, C2, cout, cout_pad, stride, act, planar, pitch, out_c,
                                     _lib.cur_stream())
    _lib.check(rc, "faln_conv3x3_fwd")
    return y
