"""Layout conversions between the reference's fp32 NCHW tensors and bf16 NHWC activations (csrc/misc.cu)."""
from __future__ import annotations

import torch

from . import _lib


def nchw_to_nhwc_bf16(x, Cp, flip_x=False):
    B, C, H, W = x.shape
    x = _lib.f32c(x)
    out = torch.empty(B, H, W, Cp, device=x.device, dtype=torch.bfloat16)
    rc = _lib.lib().faln_nchw_to_nhwc_bf16(_lib.ptr(x), _lib.ptr(out), B, C, H, W, Cp, int(flip_x), _lib.cur_stream())
    _lib.check(rc, "faln_nchw_to_nhwc_bf16")
    return out


def _pitch(x):
    B, C, H, W = x.shape
    sb, sc, sh, sw = x.stride()
    assert sw == 1 and sc == H * sh and sb == C * H * sh and sh >= W, "need a uniformly pitched planar tensor"
    return sh


def planar_to_nhwc_bf16(x, Cp):
    B, C, H, W = x.shape
    out = torch.empty(B, H, W, Cp, device=x.device, dtype=torch.bfloat16)
    rc = _lib.lib().faln_planar_to_nhwc_bf16(_lib.ptr(x), _lib.ptr(out), B, C, H, W, Cp, _pitch(x), _lib.cur_stream())
    _lib.check(rc, "faln_planar_to_nhwc_bf16")
    return out


def alloc_planar(B, C, H, W, device, pitch=None):
    """fp32 [B,C,H,W] view with a 16-byte-multiple row pitch (TMA-friendly rows for the MED kernels)."""
    pitch = pitch or ((W + 3) // 4) * 4
    buf = torch.empty(B, C, H, pitch, device=device, dtype=torch.float32)
    if pitch != W:
        buf[..., W:].zero_()          # the MED fast kernels read the pad columns as zeros (FALN_MED_ZERO_PAD)
    return buf[..., :W]


def nhwc_bf16_to_planar(x, C, pitch=None):
    B, H, W, Cp = x.shape
    out = alloc_planar(B, C, H, W, x.device, pitch)
    rc = _lib.lib().faln_nhwc_bf16_to_planar(_lib.ptr(x), _lib.ptr(out), B, C, H, W, Cp, _pitch(out), _lib.cur_stream())
    _lib.check(rc, "faln_nhwc_bf16_to_planar")
    return out
