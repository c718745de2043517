"""Device-side post-processing of Test_KITTI (csrc/postproc.cu): the pieces of ``ms_pp``
(/root/reference/Test_KITTI.py:287-300) that the reference runs as ATen launches plus one ``np.percentile`` host sync per
image.  No CPU path."""
from __future__ import annotations

import math

import torch

from . import _lib


def flip_resize_bilinear(x, scale_factor=None, size=None, flip_x=True):
    """F.interpolate(flip(x), scale_factor, mode='bilinear', align_corners=True) in one kernel.  x fp32 [B,C,H,W]."""
    x = _lib.f32c(x, "image")
    B, C, H, W = x.shape
    if size is None:
        size = (int(math.floor(float(H) * scale_factor)), int(math.floor(float(W) * scale_factor)))   # ATen's output size
    Ho, Wo = size
    out = torch.empty(B, C, Ho, Wo, device=x.device, dtype=torch.float32)
    _lib.check(_lib.lib().faln_flip_resize_bilinear(_lib.ptr(x), _lib.ptr(out), B * C, H, W, Ho, Wo, int(flip_x),
                                                    _lib.cur_stream()), "faln_flip_resize_bilinear")
    return out


def percentile_rows(x, q, add=0.0):
    """[B] fp32: numpy.percentile(x[b].ravel(), q) + add for every sample b, on the device, exact (no host sync)."""
    x = _lib.f32c(x, "x")
    B = x.shape[0]
    n = x.numel() // B
    out = torch.empty(B, device=x.device, dtype=torch.float32)
    _lib.check(_lib.lib().faln_percentile_rows(_lib.ptr(x), B, n, n, float(q) / 100.0, float(add), _lib.ptr(out),
                                               _lib.cur_stream()), "faln_percentile_rows")
    return out


def mspp_blend(disp, small_flipped, p, up_mul):
    """(1 - norm) * disp + norm * up_mul * unflip(nearest_up(small)), norm = min(disp / p[b], 1)."""
    disp, small = _lib.f32c(disp, "disp"), _lib.f32c(small_flipped, "small")
    B, _, H, W = disp.shape
    Hs, Ws = small.shape[2], small.shape[3]
    out = torch.empty_like(disp)
    _lib.check(_lib.lib().faln_mspp_blend(_lib.ptr(disp), _lib.ptr(small), _lib.ptr(p), _lib.ptr(out), B, H, W, Hs, Ws,
                                          float(up_mul), _lib.cur_stream()), "faln_mspp_blend")
    return out
