"""MED view synthesis on the fused sm_100a kernels (csrc/med.cu).

Host-side mirror of /root/reference/models/FAL_netB.py:200-297 from ``dlog0`` onwards: the level
tables are computed with the reference's own torch expressions (so their fp32 values are the
reference's), everything per-pixel happens in ``faln_med_fwd`` / ``faln_med_bwd``.
"""
from __future__ import annotations

import torch
import torch.nn.functional as F

from . import _lib


def level_tables(min_disp: torch.Tensor, max_disp: torch.Tensor, no_levels: int, W: int):
    """d [B,N] (pixels) and x_of [B,N] (normalised grid units) of the exponential disparity levels: the fp32 expressions of
    /root/reference/models/FAL_netB.py:204-205,224-225,241 for all levels in ONE launch (``faln_level_tables`` replays the
    reference's op order with every operation rounded to fp32; the reference launches ~6 tiny ATen kernels per level per
    loop, ~300 per forward).  Bit-compared with the torch expressions on the device in tests/test_med_gpu.py."""
    B = min_disp.shape[0]
    mn = _lib.f32c(min_disp.reshape(B), "min_disp")
    mx = _lib.f32c(max_disp.reshape(B), "max_disp")
    d = torch.empty(B, no_levels, device=mn.device, dtype=torch.float32)
    xo = torch.empty(B, no_levels, device=mn.device, dtype=torch.float32)
    _lib.check(_lib.lib().faln_level_tables(_lib.ptr(mn), _lib.ptr(mx), _lib.ptr(d), _lib.ptr(xo), B, no_levels, W,
                                            _lib.cur_stream()), "faln_level_tables")
    return d, xo


def level_tables_torch(min_disp: torch.Tensor, max_disp: torch.Tensor, no_levels: int, W: int):
    """The same tables with the reference's own torch expressions, vectorised over the levels (the per-level factor (c - 1),
    a Python double in the reference, reaches the fp32 kernel rounded to fp32).  Test reference for ``level_tables``."""
    cm1 = torch.tensor([n / (no_levels - 1) - 1 for n in range(no_levels)], dtype=torch.float64).to(torch.float32)
    cm1 = cm1.to(min_disp.device)
    x_pix_min = 2 * min_disp / W
    x_pix_max = 2 * max_disp / W
    d = max_disp * torch.exp(torch.log(max_disp / min_disp) * cm1)
    xo = x_pix_max * torch.exp(torch.log(x_pix_max / x_pix_min) * cm1)
    return d.squeeze(1).contiguous(), xo.squeeze(1).contiguous()


_G0X_CACHE: dict = {}

# bench.py sets this to a list to collect (kind, start_event, end_event, algorithmic_bytes) per MED launch
TIMING = None


def _timed(kind, nbytes):
    if TIMING is None:
        return None
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    TIMING.append((kind, e0, e1, nbytes))
    e0.record()
    return e1


def grid_row(W: int, device) -> torch.Tensor:
    """x row of F.affine_grid(identity, align_corners=True) (/root/reference/models/FAL_netB.py:231-234)."""
    key = (W, str(device))
    g = _G0X_CACHE.get(key)
    if g is None:
        th = torch.zeros(1, 2, 3, device=device)
        th[:, 0, 0] = 1
        th[:, 1, 1] = 1
        g = F.affine_grid(th, [1, 1, 2, W], align_corners=True)[0, 0, :, 0].contiguous()
        _G0X_CACHE[key] = g
    return g


def _pitch_of(logits: torch.Tensor) -> int:
    B, N, H, W = logits.shape
    sb, sn, sh, sw = logits.stride()
    if sw != 1 or sn != H * sh or sb != N * H * sh or sh < W:
        raise RuntimeError("logits must be [B,N,H,W] with unit x stride and a uniform row pitch")
    return sh


def med_forward_raw(logits, image, x_of, d_lvl, g0x, want_pan=True, want_disp=True, want_masks=False, flags=0):
    """Direct kernel call.  Returns dict(pan, disp, maskL, maskR, lse0, lsew) (None where not wanted)."""
    L = _lib.lib()
    B, N, H, W = logits.shape
    assert logits.dtype == torch.float32 and logits.is_cuda
    pitch = _pitch_of(logits)
    image = _lib.f32c(image, "image")
    assert image.shape == (B, 3, H, W)
    dev = logits.device
    opt = dict(device=dev, dtype=torch.float32)
    pan = torch.empty(B, 3, H, W, **opt) if want_pan else None
    disp = torch.empty(B, 1, H, W, **opt) if want_disp else None
    mL = torch.empty(B, 1, H, W, **opt) if want_masks else None
    mR = torch.empty(B, 1, H, W, **opt) if want_masks else None
    lse0 = torch.empty(B, 1, H, W, **opt)
    lsew = torch.empty(B, 1, H, W, **opt)
    # algorithmic bytes: read N logits + 3 image, write 3 pan + 1 disp (+ 2 masks) = 4(N+9) / 4(N+7)
    # B/px; the two saved log-sum-exp rows (8 B/px) are extra traffic of this design and are NOT counted
    ev = _timed("med_fwd_masks" if want_masks else "med_fwd", 4 * (N + 7 + (2 if want_masks else 0)) * B * H * W)
    rc = L.faln_med_fwd(_lib.ptr(logits), _lib.ptr(image), _lib.ptr(g0x), _lib.ptr(x_of), _lib.ptr(d_lvl),
                        _lib.ptr(pan), _lib.ptr(disp), _lib.ptr(mL), _lib.ptr(mR), _lib.ptr(lse0), _lib.ptr(lsew),
                        B, N, H, W, pitch, flags, _lib.cur_stream())
    if ev is not None:
        ev.record()
    _lib.check(rc, "faln_med_fwd")
    return dict(pan=pan, disp=disp, maskL=mL, maskR=mR, lse0=lse0, lsew=lsew)


def med_backward_raw(logits, image, x_of, d_lvl, g0x, pan, disp, lse0, lsew, g_pan, g_disp, flags=0, out=None):
    L = _lib.lib()
    B, N, H, W = logits.shape
    pitch = _pitch_of(logits)
    if out is None:
        out = torch.empty_like(logits)
    g_pitch = _pitch_of(out)
    g_pan = _lib.f32c(g_pan, "g_pan") if g_pan is not None else None
    g_disp = _lib.f32c(g_disp, "g_disp") if g_disp is not None else None
    ev = _timed("med_bwd", 4 * (2 * N + 7) * B * H * W)
    rc = L.faln_med_bwd(_lib.ptr(logits), _lib.ptr(image), _lib.ptr(g0x), _lib.ptr(x_of), _lib.ptr(d_lvl),
                        _lib.ptr(pan), _lib.ptr(disp), _lib.ptr(lse0), _lib.ptr(lsew), _lib.ptr(g_pan),
                        _lib.ptr(g_disp), _lib.ptr(out), B, N, H, W, pitch, g_pitch, flags, _lib.cur_stream())
    if ev is not None:
        ev.record()
    _lib.check(rc, "faln_med_bwd")
    return out


def med_disp_only(logits, d_lvl):
    """Inference epilogue: disparity expectation only (/root/reference/models/FAL_netB.py:216-229)."""
    L = _lib.lib()
    B, N, H, W = logits.shape
    disp = torch.empty(B, 1, H, W, device=logits.device, dtype=torch.float32)
    rc = L.faln_med_disp(_lib.ptr(logits), _lib.ptr(d_lvl), _lib.ptr(disp), B, N, H, W, _pitch_of(logits),
                         _lib.cur_stream())
    _lib.check(rc, "faln_med_disp")
    return disp


class MedSynthesis(torch.autograd.Function):
    """(logits, image) -> (pan, disp, maskL, maskR); gradient flows to the logits only, from pan and disp
    only -- exactly the reference's autograd graph (masks are built under no_grad on detached inputs,
    /root/reference/models/FAL_netB.py:256-273; the image is a leaf without grad)."""

    @staticmethod
    def forward(ctx, logits, image, x_of, d_lvl, g0x, want_masks, flags=0):
        r = med_forward_raw(logits, image, x_of, d_lvl, g0x, True, True, want_masks, flags)
        ctx.flags = flags
        ctx.save_for_backward(logits, image, x_of, d_lvl, g0x, r["pan"], r["disp"], r["lse0"], r["lsew"])
        outs = (r["pan"], r["disp"])
        if want_masks:
            ctx.mark_non_differentiable(r["maskL"], r["maskR"], r["lse0"])
            outs = outs + (r["maskL"], r["maskR"], r["lse0"])
        return outs

    @staticmethod
    def backward(ctx, g_pan, g_disp, *_):
        logits, image, x_of, d_lvl, g0x, pan, disp, lse0, lsew = ctx.saved_tensors
        g = med_backward_raw(logits, image, x_of, d_lvl, g0x, pan, disp, lse0, lsew, g_pan, g_disp, ctx.flags)
        return g, None, None, None, None, None, None


FLAG_ZERO_PAD = 8     # FALN_MED_ZERO_PAD
FLAG_NO_FAST = 16     # FALN_MED_NO_FAST
FLAG_NO_V3 = 32       # FALN_MED_NO_V3: second-generation kernels (A/B comparison)
FLAG_V3_GENERIC = 64  # FALN_MED_V3_GENERIC: third generation, every plane on the per-pixel generic code (testing)


def maskr_noalign(logits, lse0, x_of):
    """FAL_netA's maskR (/root/reference/models/FAL_netA.py:264: grid_sample with its default align_corners=False)."""
    B, N, H, W = logits.shape
    out = torch.empty(B, 1, H, W, device=logits.device, dtype=torch.float32)
    rc = _lib.lib().faln_maskr_noalign(_lib.ptr(logits), _lib.ptr(lse0), _lib.ptr(grid_row(W, logits.device)),
                                       _lib.ptr(grid_row(H, logits.device)), _lib.ptr(x_of), _lib.ptr(out), B, N, H, W,
                                       _pitch_of(logits), _lib.cur_stream())
    _lib.check(rc, "faln_maskr_noalign")
    return out


def med_section(dlog0, image, min_disp, max_disp, ret_disp=True, ret_subocc=False, ret_pan=False, zero_pad=False,
                maskr_align_corners=True):
    """FAL_net.forward from ``dlog0`` on, same return convention as the reference
    (/root/reference/models/FAL_netB.py:228-229,285-297): a bare tensor when only the disparity is asked
    for, else a list ordered [pan?, disp?, maskL?, maskR?].  ``zero_pad``: the caller guarantees that ``dlog0`` comes from
    ``layout.alloc_planar`` (16-byte aligned rows, zeroed pad columns), which enables the fast kernels for W % 4 != 0."""
    B, N, H, W = dlog0.shape
    flags = FLAG_ZERO_PAD if zero_pad else 0
    d_lvl, x_of = level_tables(min_disp, max_disp, N, W)
    if ret_disp and not ret_subocc and not ret_pan:
        if dlog0.requires_grad and torch.is_grad_enabled():
            g0x = grid_row(W, dlog0.device)
            return MedSynthesis.apply(dlog0, image, x_of, d_lvl, g0x, False, flags)[1]
        return med_disp_only(dlog0, d_lvl)
    g0x = grid_row(W, dlog0.device)
    res = MedSynthesis.apply(dlog0, image, x_of, d_lvl, g0x, bool(ret_subocc), flags)
    if ret_subocc and not maskr_align_corners:
        # FAL_netA: maskR is re-sampled with align_corners=False; the fused kernel's pan / disp / maskL stand
        with torch.no_grad():
            res = res[:3] + (maskr_noalign(dlog0.detach(), res[4], x_of),)
    out = []
    if ret_pan:
        out.append(res[0])
    if ret_disp:
        out.append(res[1])
    if ret_subocc:
        out.extend([res[2], res[3]])
    return out
