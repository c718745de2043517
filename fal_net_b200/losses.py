"""Loss kernels of libfalnet_sm100.so behind thin torch wrappers (csrc/losses.cu).

Raw functions (``rec_l1`` ...) enqueue one kernel each and never synchronise; ``loss_functions.py`` at
the package root composes them into the reference's ``rec_loss_fnc`` / ``smoothness`` signatures.
"""
from __future__ import annotations

import torch

from . import _lib

_WS: dict = {}


def _workspace(dev) -> torch.Tensor:
    """Reduction scratch: a zero-initialised ticket counter + per-block partials.  One per (device,
    stream): kernels on one stream are ordered, so sharing it between consecutive loss calls is safe."""
    key = (str(dev), torch.cuda.current_stream(dev).cuda_stream)
    ws = _WS.get(key)
    if ws is None:
        ws = torch.zeros(_lib.lib().faln_loss_partials_len(), device=dev, dtype=torch.float32)
        _WS[key] = ws
    return ws


def _scalar(dev):
    return torch.empty(1, device=dev, dtype=torch.float32)


def rec_l1(synth, label, mask=None, want_blend=False, flip_x=False):
    B, C, H, W = label.shape
    assert C == 3 and synth.shape == label.shape
    synth, label = _lib.f32c(synth), _lib.f32c(label)
    mask = _lib.f32c(mask) if mask is not None else None
    out = _scalar(label.device)
    blend = torch.empty_like(label) if want_blend else None
    rc = _lib.lib().faln_loss_rec_l1(_lib.ptr(synth), _lib.ptr(label), _lib.ptr(mask), _lib.ptr(blend), _lib.ptr(out),
                                     _lib.ptr(_workspace(label.device)), B, H, W, int(flip_x), _lib.cur_stream())
    _lib.check(rc, "faln_loss_rec_l1")
    return (out[0], blend) if want_blend else out[0]


def rec_l1_bwd(synth, label, mask, g_blend, g_scale, flip_x=False, g_dev=None):
    B, C, H, W = label.shape
    synth, label = _lib.f32c(synth), _lib.f32c(label)
    mask = _lib.f32c(mask) if mask is not None else None
    g_blend = _lib.f32c(g_blend) if g_blend is not None else None
    g = torch.empty_like(synth)
    rc = _lib.lib().faln_loss_rec_l1_bwd(_lib.ptr(synth), _lib.ptr(label), _lib.ptr(mask), _lib.ptr(g_blend),
                                         float(g_scale), _lib.ptr(g_dev), _lib.ptr(g), B, H, W, int(flip_x),
                                         _lib.cur_stream())
    _lib.check(rc, "faln_loss_rec_l1_bwd")
    return g


def smoothness(img, disp, gamma, x_lo, x_hi, flip_x=False):
    B, _, H, W = img.shape
    img, disp = _lib.f32c(img), _lib.f32c(disp)
    out = _scalar(img.device)
    rc = _lib.lib().faln_loss_smooth(_lib.ptr(img), _lib.ptr(disp), float(gamma), _lib.ptr(out),
                                     _lib.ptr(_workspace(img.device)), B, H, W, x_lo, x_hi, int(flip_x),
                                     _lib.cur_stream())
    _lib.check(rc, "faln_loss_smooth")
    return out[0]


def smoothness_bwd(img, disp, gamma, g_scale, g_disp, x_lo, x_hi, flip_x=False, g_dev=None, accumulate=True):
    B, _, H, W = img.shape
    img, disp = _lib.f32c(img), _lib.f32c(disp)
    assert g_disp.is_contiguous() and g_disp.dtype == torch.float32
    rc = _lib.lib().faln_loss_smooth_bwd(_lib.ptr(img), _lib.ptr(disp), float(gamma), float(g_scale), _lib.ptr(g_dev),
                                         _lib.ptr(g_disp), int(accumulate), B, H, W, x_lo, x_hi, int(flip_x),
                                         _lib.cur_stream())
    _lib.check(rc, "faln_loss_smooth_bwd")
    return g_disp


def mirror(disp, mdisp, occ, inv_max, x_lo, x_hi, flip_x=False):
    B, _, H, W = disp.shape
    out = _scalar(disp.device)
    rc = _lib.lib().faln_loss_mirror(_lib.ptr(_lib.f32c(disp)), _lib.ptr(_lib.f32c(mdisp)), _lib.ptr(_lib.f32c(occ)),
                                     _lib.ptr(inv_max), _lib.ptr(out), _lib.ptr(_workspace(disp.device)), B, H, W, x_lo,
                                     x_hi, int(flip_x), _lib.cur_stream())
    _lib.check(rc, "faln_loss_mirror")
    return out[0]


def mirror_bwd(disp, mdisp, occ, inv_max, g_scale, g_disp, x_lo, x_hi, flip_x=False, g_dev=None, accumulate=True):
    B, _, H, W = disp.shape
    rc = _lib.lib().faln_loss_mirror_bwd(_lib.ptr(_lib.f32c(disp)), _lib.ptr(_lib.f32c(mdisp)), _lib.ptr(_lib.f32c(occ)),
                                         _lib.ptr(inv_max), float(g_scale), _lib.ptr(g_dev), _lib.ptr(g_disp),
                                         int(accumulate), B, H, W, x_lo, x_hi, int(flip_x), _lib.cur_stream())
    _lib.check(rc, "faln_loss_mirror_bwd")
    return g_disp


def mse_bf16(a, b):
    assert a.dtype == torch.bfloat16 and b.dtype == torch.bfloat16 and a.is_contiguous() and b.is_contiguous()
    out = _scalar(a.device)
    rc = _lib.lib().faln_mse_bf16(_lib.ptr(a), _lib.ptr(b), a.numel(), _lib.ptr(out), _lib.ptr(_workspace(a.device)),
                                  _lib.cur_stream())
    _lib.check(rc, "faln_mse_bf16")
    return out[0]


def mse_bf16_bwd(a, b, g_scale, g_dev=None):
    g = torch.empty_like(a)
    rc = _lib.lib().faln_mse_bf16_bwd(_lib.ptr(a), _lib.ptr(b), a.numel(), float(g_scale), _lib.ptr(g_dev), _lib.ptr(g),
                                      _lib.cur_stream())
    _lib.check(rc, "faln_mse_bf16_bwd")
    return g


def inv_rowmax(x):
    x = _lib.f32c(x)
    B = x.shape[0]
    out = torch.empty(B, device=x.device, dtype=torch.float32)
    rc = _lib.lib().faln_inv_rowmax(_lib.ptr(x), _lib.ptr(out), B, x.numel() // B, _lib.cur_stream())
    _lib.check(rc, "faln_inv_rowmax")
    return out


def occ_mask(a, b, flip_a, flip_b, one_lo, one_hi):
    B, _, H, W = a.shape
    a, b = _lib.f32c(a), _lib.f32c(b)
    out = torch.empty_like(a)
    rc = _lib.lib().faln_occ_mask(_lib.ptr(a), _lib.ptr(b), _lib.ptr(out), B, H, W, int(flip_a), int(flip_b), one_lo,
                                  one_hi, _lib.cur_stream())
    _lib.check(rc, "faln_occ_mask")
    return out
