"""Fused Adam over a flat parameter arena (csrc/misc.cu: adam_kernel)."""
from __future__ import annotations

import torch

from . import _lib


def adam_step_(p, g, m, v, w16=None, *, lr, beta1=0.5, beta2=0.999, eps=1e-8, weight_decay=0.0, step=1, grad_scale=1.0):
    """In-place torch.optim.Adam update of the flat fp32 arena ``p`` (numel % 4 == 0) with gradient ``g``;
    also refreshes the bf16 shadow ``w16`` used by the conv kernels.  Defaults are the reference's
    (/root/reference/Train_Stage1_K.py:177-181)."""
    n = p.numel()
    rc = _lib.lib().faln_adam(_lib.ptr(p), _lib.ptr(g), _lib.ptr(m), _lib.ptr(v), _lib.ptr(w16), n, float(lr),
                              float(beta1), float(beta2), float(eps), float(weight_decay), int(step), float(grad_scale),
                              _lib.cur_stream())
    _lib.check(rc, "faln_adam")


def adam_step_dev_(p, g, m, v, hp, w16=None, *, beta1=0.5, beta2=0.999, eps=1e-8, weight_decay=0.0, grad_scale=1.0):
    """Same update with the learning rate and step counter in device memory (``hp`` = float32[4]: lr, step, and two
    derived factors the kernel maintains) -- no host scalar changes between steps, so the call can be captured in a
    CUDA graph and replayed."""
    rc = _lib.lib().faln_adam_dev(_lib.ptr(p), _lib.ptr(g), _lib.ptr(m), _lib.ptr(v), _lib.ptr(w16), p.numel(),
                                  _lib.ptr(hp), float(beta1), float(beta2), float(eps), float(weight_decay),
                                  float(grad_scale), _lib.cur_stream())
    _lib.check(rc, "faln_adam_dev")


def adam_range_dev_(p, g, m, v, hp, w16=None, *, tick, beta1=0.5, beta2=0.999, eps=1e-8, weight_decay=0.0, grad_scale=1.0):
    """``adam_step_dev_`` over one contiguous range of the arenas (all arguments are the range's slices); ``tick`` advances
    the device-side step counter first -- the first range of a step ticks, the others do not."""
    rc = _lib.lib().faln_adam_dev_range(_lib.ptr(p), _lib.ptr(g), _lib.ptr(m), _lib.ptr(v), _lib.ptr(w16), p.numel(),
                                        _lib.ptr(hp), float(beta1), float(beta2), float(eps), float(weight_decay),
                                        float(grad_scale), 1 if tick else 0, _lib.cur_stream())
    _lib.check(rc, "faln_adam_dev_range")
