"""Thin torch wrappers over the tcgen05 implicit-GEMM convolution kernels (csrc/conv_tc.cu, csrc/conv_aux.cu).

Activations: bf16 tensors of logical shape [B,C,H,W] in channels_last memory format (= NHWC in memory)."""
from __future__ import annotations

import ctypes

import torch

from . import _lib

CL = torch.channels_last
ACT = {None: 0, "none": 0, "elu": 1, "relu": 2}

# bench.py sets this to a list to collect (kind, start_event, end_event, algorithmic_flops, algorithmic_bytes) per conv launch
TIMING = None


def _timed(kind, flops, nbytes, gemm_n=64):
    if TIMING is None:
        return None
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    TIMING.append((kind, e0, e1, flops, nbytes, gemm_n))
    e0.record()
    return e1


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


def const_channel_table_of(weight: torch.Tensor, channel: int) -> torch.Tensor:
    """[16,Cout] border-class sums of input channel ``channel`` of a (possibly KRSC-strided) [Cout,Cin,3,3] fp32 weight:
    one launch (csrc/small_ops.cu), no slicing / matmul."""
    w = weight.detach()
    assert w.dtype == torch.float32 and w.dim() == 4 and w.shape[2:] == (3, 3)
    ctab = torch.empty(16, w.shape[0], device=w.device, dtype=torch.float32)
    so, sc, sh, sw = w.stride()
    _lib.check(_lib.lib().faln_const_channel_table(_lib.ptr(w), so, sc, sh, sw, int(channel), _lib.ptr(ctab), w.shape[0],
                                                   _lib.cur_stream()), "faln_const_channel_table")
    return ctab


def fold_logit_conv(w_iconv1: torch.Tensor, w0: torch.Tensor, rows_pad=None):
    """iconv1 (3x3, no bias, no activation; reference :127,174) followed by conv0 (1x1 + bias; :190,215) == one 3x3 conv
    with W'[o,c,kh,kw] = sum_m W0[o,m] * W_iconv1[m,c,kh,kw].  Returns (forward pack [Np,3,3,C] bf16, data-gradient pack
    [Cp,3,3,Np] bf16) straight from one kernel (csrc/small_ops.cu)."""
    wi, w0 = w_iconv1.detach(), w0.detach()
    N, C = wi.shape[0], wi.shape[1]
    Np = rows_pad or (N + 31) // 32 * 32
    Cp = (C + 31) // 32 * 32
    fwd = torch.empty(Np, 3, 3, C, device=wi.device, dtype=torch.bfloat16)
    dg = torch.empty(Cp, 3, 3, Np, device=wi.device, dtype=torch.bfloat16)
    so, sc, sh, sw = wi.stride()
    w0m = w0.reshape(N, N)
    assert w0m.is_contiguous() and wi.dtype == torch.float32 and w0m.dtype == torch.float32
    _lib.check(_lib.lib().faln_fold_logit_conv(_lib.ptr(w0m), _lib.ptr(wi), so, sc, sh, sw, _lib.ptr(fwd), _lib.ptr(dg), N, C,
                                               Np, Cp, _lib.cur_stream()), "faln_fold_logit_conv")
    return fwd, dg


def fold_logit_conv_bwd(gwf, w_iconv1, w0, g_w_iconv1, g_w0):
    """Adjoint of ``fold_logit_conv``: accumulates into the gradient views g_w_iconv1 (strided) and g_w0 ([N,N,1,1])."""
    wi, w0 = w_iconv1.detach(), w0.detach()
    N, C = wi.shape[0], wi.shape[1]
    gk = gwf.permute(0, 2, 3, 1)
    assert gk.is_contiguous() and gk.shape == (N, 3, 3, C) and g_w0.is_contiguous()
    so, sc, sh, sw = wi.stride()
    gso, gsc, gsh, gsw = g_w_iconv1.stride()
    _lib.check(_lib.lib().faln_fold_logit_conv_bwd(_lib.ptr(gk), _lib.ptr(w0.reshape(N, N)), _lib.ptr(wi), so, sc, sh, sw,
                                                   _lib.ptr(g_w_iconv1), gso, gsc, gsh, gsw, _lib.ptr(g_w0), N, C,
                                                   _lib.cur_stream()), "faln_fold_logit_conv_bwd")


def const_channel_wgrad_into(g, value, in_hw, stride, cout, dW, channel):
    """dW[:, channel] += weight gradient of a spatially constant input channel with per-sample value ``value`` [B]: the
    nine border-class sums of the output gradient (one kernel) combined per tap (one kernel)."""
    B, Cs, Hg, Wg = g.shape
    H, W = in_hw
    S = border_sums(g, cout)                                               # [B,3,3,C]
    so, sc, sh, sw = dW.stride()
    val = value.float().contiguous()
    _lib.check(_lib.lib().faln_const_channel_wgrad(_lib.ptr(S), _lib.ptr(val), _lib.ptr(dW), so, sc, sh, sw, int(channel), B,
                                                   cout, int(stride * (Hg - 1) + 1 > H - 1), int(stride * (Wg - 1) + 1 > W - 1),
                                                   _lib.cur_stream()), "faln_const_channel_wgrad")


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
    ev = _timed("conv_fwd", 2 * 9 * (C1 + C2) * cout * B * Ho * Wo,
                2 * B * H * W * (C1 + C2) + (4 if planar else 2) * B * Ho * Wo * cout + 2 * 9 * (C1 + C2) * cout, cout_pad)
    rc = _lib.lib().faln_conv3x3_fwd(_lib.ptr(x), _lib.ptr(x2), _lib.ptr(w_krsc), _lib.ptr(bias), _lib.ptr(ctab),
                                     _lib.ptr(cscale), _lib.ptr(residual), _lib.ptr(y), B, H, W, C1, C2, cout, cout_pad,
                                     stride, int(act), planar, pitch, out_c, _lib.cur_stream())
    _lib.check(rc, "faln_conv3x3_fwd")
    if ev is not None:
        ev.record()
    return y


def logits_disp_supported(W, n_levels):
    """The fused logits -> disparity epilogue lives in the row-tile kernel: wide maps, at most 64 levels."""
    return W >= 192 and n_levels <= 64


def conv3x3_logits_disp(x, x2, w_krsc, bias, d_lvl):
    """Inference form of the last layer: disparity [B,1,H,W] fp32 = sum_n d_lvl[b,n] * softmax_n(conv3x3(cat(x, x2)) + bias),
    with the N logit planes kept on chip (csrc/conv_tc.cu, row-tile kernel)."""
    x, x2 = _nhwc(x), _nhwc(x2)
    B, C1, H, W = x.shape
    C2 = x2.shape[1]
    N = d_lvl.shape[1]
    assert w_krsc.shape == (64, 3, 3, C1 + C2) and w_krsc.dtype == torch.bfloat16 and w_krsc.is_contiguous()
    assert d_lvl.shape == (B, N) and N <= 64
    d64 = torch.zeros(B, 64, device=x.device, dtype=torch.float32)          # levels beyond N: weight 0 ...
    d64[:, :N] = d_lvl
    b64 = torch.full((64,), float("-inf"), device=x.device, dtype=torch.float32)   # ... and logit -inf
    b64[:N] = bias.detach().float()
    disp = torch.empty(B, 1, H, W, device=x.device, dtype=torch.float32)
    ev = _timed("conv_fwd", 2 * 9 * (C1 + C2) * N * B * H * W, 2 * B * H * W * (C1 + C2) + 4 * B * H * W, 64)
    rc = _lib.lib().faln_conv3x3_logits_disp(_lib.ptr(x), _lib.ptr(x2), _lib.ptr(w_krsc), _lib.ptr(b64), _lib.ptr(d64),
                                             _lib.ptr(disp), B, H, W, C1, C2, N, 64, _lib.cur_stream())
    _lib.check(rc, "faln_conv3x3_logits_disp")
    if ev is not None:
        ev.record()
    return disp


def pack_up2_weights(weight: torch.Tensor):
    """(forward pack [Cout_pad,16,Cin] bf16, dgrad pack [Cin_pad,16,Cout_pad] bf16) of an up-sample + conv block's folded
    weights (csrc/small_ops.cu pack_up2_kernel), from the fp32 [Cout,Cin,3,3] weight in whatever layout it lives."""
    w = weight.detach()
    assert w.dtype == torch.float32 and w.dim() == 4 and w.shape[2:] == (3, 3)
    Cout, Cin = w.shape[0], w.shape[1]
    Cout_pad, Cin_pad = (Cout + 31) // 32 * 32, (Cin + 31) // 32 * 32
    assert Cin % 32 == 0
    fwd = torch.empty(Cout_pad, 16, Cin, device=w.device, dtype=torch.bfloat16)
    dg = torch.empty(Cin_pad, 16, Cout_pad, device=w.device, dtype=torch.bfloat16)
    so, sc, sh, sw = w.stride()
    _lib.check(_lib.lib().faln_pack_up2_weights(_lib.ptr(w), so, sc, sh, sw, _lib.ptr(fwd), _lib.ptr(dg), Cout, Cin, Cout_pad,
                                                Cin_pad, _lib.cur_stream()), "faln_pack_up2_weights")
    return fwd, dg


class _Up2PackJob(ctypes.Structure):
    _fields_ = ([("w", ctypes.c_void_p)] + [(n, ctypes.c_longlong) for n in ("so", "sc", "sh", "sw")] +
                [("fwd_pack", ctypes.c_void_p), ("dgrad_pack", ctypes.c_void_p)] +
                [(n, ctypes.c_int) for n in ("Cout", "Cin", "Cout_pad", "Cin_pad")])


def pack_up2_weights_multi(weights):
    """``pack_up2_weights`` of several deconv weights in ONE launch; returns a list of (forward pack, dgrad pack)."""
    n = len(weights)
    assert 0 < n <= 8
    arr = (_Up2PackJob * n)()
    out = []
    for i, weight in enumerate(weights):
        w = weight.detach()
        assert w.dtype == torch.float32 and w.dim() == 4 and w.shape[2:] == (3, 3) and w.is_cuda
        Cout, Cin = w.shape[0], w.shape[1]
        Cout_pad, Cin_pad = (Cout + 31) // 32 * 32, (Cin + 31) // 32 * 32
        assert Cin % 32 == 0
        fwd = torch.empty(Cout_pad, 16, Cin, device=w.device, dtype=torch.bfloat16)
        dg = torch.empty(Cin_pad, 16, Cout_pad, device=w.device, dtype=torch.bfloat16)
        a = arr[i]
        a.w = w.data_ptr()
        a.so, a.sc, a.sh, a.sw = w.stride()
        a.fwd_pack, a.dgrad_pack = fwd.data_ptr(), dg.data_ptr()
        a.Cout, a.Cin, a.Cout_pad, a.Cin_pad = Cout, Cin, Cout_pad, Cin_pad
        out.append((fwd, dg))
    _lib.check(_lib.lib().faln_pack_up2_weights_multi(ctypes.cast(arr, ctypes.c_void_p), n, _lib.cur_stream()),
               "faln_pack_up2_weights_multi")
    return out


def conv3x3_up2_fwd(x, w_fold, bias=None, act=0, cout=None):
    """act(conv3x3(upsample_nearest_2x(x))) without the up-sampled tensor: x bf16 [B,C,H,W] channels_last (low resolution),
    w_fold from ``pack_up2_weights``; returns bf16 [B,Cout,2H,2W] channels_last."""
    x = _nhwc(x)
    B, C1, H, W = x.shape
    cout_pad = w_fold.shape[0]
    cout = cout or cout_pad
    assert w_fold.shape[1:] == (16, C1) and w_fold.dtype == torch.bfloat16 and w_fold.is_contiguous()
    y = torch.empty((B, cout, 2 * H, 2 * W), device=x.device, dtype=torch.bfloat16, memory_format=CL)
    if bias is not None:
        bias = bias.detach().float().contiguous()
    # algorithmic FLOPs of the REFERENCE formulation (nine taps on the up-sampled map) so that rooflines stay comparable
    ev = _timed("conv_fwd", 2 * 9 * C1 * cout * B * 4 * H * W, 2 * B * H * W * C1 + 2 * B * 4 * H * W * cout + 2 * 9 * C1 * cout,
                cout_pad)
    rc = _lib.lib().faln_conv3x3_up2_fwd(_lib.ptr(x), _lib.ptr(w_fold), _lib.ptr(bias), _lib.ptr(y), B, H, W, C1, cout, cout_pad,
                                         int(act), cout, _lib.cur_stream())
    _lib.check(rc, "faln_conv3x3_up2_fwd")
    if ev is not None:
        ev.record()
    return y


def conv3x3_up2_dgrad(g, wd_fold, dact=0, ysave=None):
    """Gradient w.r.t. the LOW-resolution input of an up-sample + conv block, times act'(ysave) of the producer:
    g bf16 [B,Cg,2H,2W] channels_last -> bf16 [B,Cx,H,W] channels_last."""
    g = _nhwc(g)
    B, Cg, H2, W2 = g.shape
    assert H2 % 2 == 0 and W2 % 2 == 0
    H, W = H2 // 2, W2 // 2
    Cx = wd_fold.shape[0]
    assert wd_fold.shape[1:] == (16, Cg) and wd_fold.dtype == torch.bfloat16 and wd_fold.is_contiguous()
    out = torch.empty((B, Cx, H, W), device=g.device, dtype=torch.bfloat16, memory_format=CL)
    if ysave is not None:
        ysave = _nhwc(ysave)
        assert ysave.shape == out.shape
    ev = _timed("conv_dgrad", 2 * 9 * Cg * Cx * B * H2 * W2, 2 * B * H2 * W2 * Cg + 2 * B * H * W * Cx + 2 * 9 * Cg * Cx, Cx)
    rc = _lib.lib().faln_conv3x3_up2_dgrad(_lib.ptr(g), _lib.ptr(wd_fold), _lib.ptr(out), _lib.ptr(ysave), B, H, W, Cg, Cx,
                                           int(dact if ysave is not None else 0), Cx, Cx, _lib.cur_stream())
    _lib.check(rc, "faln_conv3x3_up2_dgrad")
    if ev is not None:
        ev.record()
    return out


# FALN_STEM_TC=1: tensor-core stem (one im2col pass to a 32-wide bf16 K vector + a K = 32 tcgen05 GEMM) instead of the
# fp32-FMA stem kernel.  Measured on B200 (round 2): Stage-1 step 4.345 vs 4.340 ms, Stage-2 13.95 vs 14.11 ms (the
# 64-channel VGG stem gains), Test flip-PP 7.73 vs 7.65 ms (loses) -- the extra pass over a 64 B/px patch tensor and the tile
# kernel's per-pixel stores eat what the FMA pipe saves, and it rounds the input image to bf16.  Default: off.
STEM_TC = __import__("os").environ.get("FALN_STEM_TC", "0") not in ("", "0")
# The default: the fused tcgen05 stem (csrc/conv_tc.cu stem_mma_kernel; patch rows built in shared memory as bf16 hi + lo pairs).
# FALN_STEM_FMA=1 selects the fp32-FMA kernel again.
STEM_MMA = __import__("os").environ.get("FALN_STEM_FMA", "0") in ("", "0")


def stem_conv(x, w, bias, act, flip_x=False):
    """fp32 NCHW image [B,3,H,W] -> bf16 channels_last [B,Cout,H,W]; w [Cout,3,3,3] fp32 (fp32-FMA kernel; see STEM_TC)."""
    x = _lib.f32c(x)
    B, _, H, W = x.shape
    Cout = w.shape[0]
    y = torch.empty((B, Cout, H, W), device=x.device, dtype=torch.bfloat16, memory_format=CL)
    wc = w.detach().float().contiguous()
    bc = None if bias is None else bias.detach().float().contiguous()
    if STEM_TC:
        col = torch.empty((B, H, W, 32), device=x.device, dtype=torch.bfloat16)
        wpack = torch.empty((Cout, 32), device=x.device, dtype=torch.bfloat16)
        ev = _timed("conv_fwd", 2 * 27 * Cout * B * H * W, 12 * B * H * W + 2 * B * H * W * Cout, Cout)
        rc = _lib.lib().faln_stem_conv_tc(_lib.ptr(x), _lib.ptr(wc), _lib.ptr(bc), _lib.ptr(y), _lib.ptr(col), _lib.ptr(wpack),
                                          B, H, W, Cout, int(act), int(flip_x), _lib.cur_stream())
        _lib.check(rc, "faln_stem_conv_tc")
        if ev is not None:
            ev.record()
        return y
    if STEM_MMA:
        ev = _timed("conv_fwd", 2 * 27 * Cout * B * H * W, 12 * B * H * W + 2 * B * H * W * Cout, Cout)
        rc = _lib.lib().faln_stem_conv_mma(_lib.ptr(x), _lib.ptr(wc), _lib.ptr(bc), _lib.ptr(y), B, H, W, Cout, int(act),
                                           int(flip_x), _lib.cur_stream())
        _lib.check(rc, "faln_stem_conv_mma")
        if ev is not None:
            ev.record()
        return y
    rc = _lib.lib().faln_stem_conv(_lib.ptr(x), _lib.ptr(wc), _lib.ptr(bc), _lib.ptr(y), B, H, W, Cout, int(act), int(flip_x),
                                   _lib.cur_stream())
    _lib.check(rc, "faln_stem_conv")
    return y


def stem_wgrad(image, g, dW, dbias=None):
    """dW += weight gradient of the 3 -> 32 stem conv (and dbias += its bias gradient) straight from the fp32 NCHW image
    (csrc/conv_wgrad.cu stem_wgrad_mma_kernel): image [B,3,H,W] fp32, g bf16 [B,32,H,W] channels_last (pre-activation gradient),
    dW fp32 [32,3,3,3] in channels_last (KRSC) memory, dbias fp32 [32]."""
    image = _lib.f32c(image)
    g = _nhwc(g)
    B, C, H, W = image.shape
    assert C == 3 and g.shape == (B, 32, H, W), (image.shape, g.shape)
    assert dW.dtype == torch.float32 and dW.shape == (32, 3, 3, 3) and dW.permute(0, 2, 3, 1).is_contiguous()
    assert dbias is None or (dbias.dtype == torch.float32 and dbias.is_contiguous() and dbias.numel() >= 32)
    ev = _timed("conv_wgrad", 2 * 27 * 32 * B * H * W, 12 * B * H * W + 2 * B * H * W * 32, 32)
    rc = _lib.lib().faln_stem_wgrad(_lib.ptr(image), _lib.ptr(g), _lib.ptr(dW), _lib.ptr(dbias), B, H, W, _lib.cur_stream())
    _lib.check(rc, "faln_stem_wgrad")
    if ev is not None:
        ev.record()
    return dW


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


# ---------------------------------------------------------------------------------------------------------------
# backward kernels
# ---------------------------------------------------------------------------------------------------------------
def pack_weight_dgrad(w: torch.Tensor, cin_pad: int | None = None, cout_pad: int | None = None) -> torch.Tensor:
    """[Cout,Cin,3,3] -> bf16 [Cin_pad,3,3,Cout_pad]: one row per INPUT channel (the dgrad GEMM's N dimension), the conv's
    output channels along K (zero-padded to a multiple of 32, matching the zero-padded gradient tensor)."""
    Cout, Cin = w.shape[0], w.shape[1]
    cin_pad = cin_pad or (Cin + 31) // 32 * 32
    cout_pad = cout_pad or (Cout + 31) // 32 * 32
    out = torch.zeros(cin_pad, 3, 3, cout_pad, device=w.device, dtype=torch.bfloat16)
    out[:Cin, :, :, :Cout] = w.detach().permute(1, 2, 3, 0).to(torch.bfloat16)
    return out


def conv3x3_dgrad(g, wd, out_hw, stride=1, out=None, rows=None, accum=False, dact=0, ysave=None, residual=None):
    """Data gradient on the tcgen05 kernel.  g: bf16 [B,Cg,Hg,Wg] channels_last (gradient w.r.t. the conv's pre-activation,
    Cg = wd.shape[3]); wd from ``pack_weight_dgrad``; ``rows`` = (first, count) selects the input-channel range (one call per
    concatenated source); result (or ``out``, optionally accumulated into) is bf16 [B,count,H,W] channels_last with the
    fused epilogue  out = (out_old + dgrad + residual) * act'(ysave)."""
    g = _nhwc(g)
    B, Cg, Hg, Wg = g.shape
    H, W = out_hw
    assert wd.dtype == torch.bfloat16 and wd.is_contiguous() and wd.shape[3] == Cg, (wd.shape, Cg)
    first, count = rows if rows is not None else (0, wd.shape[0])
    assert first % 32 == 0 and count % 32 == 0 and first + count <= wd.shape[0]
    wslice = wd[first:first + count]
    if out is None:
        assert not accum
        out = torch.empty((B, count, H, W), device=g.device, dtype=torch.bfloat16, memory_format=CL)
    else:
        assert out.shape == (B, count, H, W) and out.dtype == torch.bfloat16 and out.is_contiguous(memory_format=CL)
    if ysave is not None:
        ysave = _nhwc(ysave)
        assert ysave.shape == out.shape
    if residual is not None:
        residual = _nhwc(residual)
        assert residual.shape == out.shape
    assert (Hg, Wg) == ((H - 1) // stride + 1, (W - 1) // stride + 1)
    ev = _timed("conv_dgrad", 2 * 9 * Cg * count * B * Hg * Wg, 2 * B * Hg * Wg * Cg + 2 * B * H * W * count + 2 * 9 * Cg * count, count)
    rc = _lib.lib().faln_conv3x3_dgrad(_lib.ptr(g), _lib.ptr(wslice), _lib.ptr(out), _lib.ptr(residual), _lib.ptr(ysave),
                                       B, H, W, Cg, count, count, stride, int(accum), int(dact if ysave is not None else 0),
                                       count, count, count, _lib.cur_stream())
    _lib.check(rc, "faln_conv3x3_dgrad")
    if ev is not None:
        ev.record()
    return out


def upsample_nearest_bwd(g_hi, lo_hw, ysave=None, dact=0, out=None, accum=False):
    """Backward of ``upsample_nearest`` fused with act'(ysave): bf16 [B,C,Hh,Wh] -> [B,C,Hl,Wl] (channels_last)."""
    g_hi = _nhwc(g_hi)
    B, C, Hh, Wh = g_hi.shape
    Hl, Wl = lo_hw
    if out is None:
        assert not accum
        out = torch.empty((B, C, Hl, Wl), device=g_hi.device, dtype=torch.bfloat16, memory_format=CL)
    if ysave is not None:
        ysave = _nhwc(ysave)
        assert ysave.shape == out.shape
    rc = _lib.lib().faln_upsample_nearest_bwd_nhwc(_lib.ptr(g_hi), _lib.ptr(ysave), _lib.ptr(out), B, Hl, Wl, Hh, Wh, C,
                                                   int(dact if ysave is not None else 0), int(accum), _lib.cur_stream())
    _lib.check(rc, "faln_upsample_nearest_bwd_nhwc")
    return out


def maxpool2_bwd(x, g_y, dact=0):
    x, g_y = _nhwc(x), _nhwc(g_y)
    B, C, Hi, Wi = x.shape
    assert g_y.shape == (B, C, Hi // 2, Wi // 2)
    g_x = torch.empty_like(x)
    rc = _lib.lib().faln_maxpool2_bwd_nhwc(_lib.ptr(x), _lib.ptr(g_y), _lib.ptr(g_x), B, Hi, Wi, C, int(dact), _lib.cur_stream())
    _lib.check(rc, "faln_maxpool2_bwd_nhwc")
    return g_x


def channel_sum(g, out, C=None):
    """out[:C] += per-channel sums of the bf16 channels_last gradient g (fp32 accumulate; ``out`` is fp32, typically a
    bias-gradient view of the flat gradient arena)."""
    g = _nhwc(g)
    B, Cs, H, W = g.shape
    C = C or Cs
    assert out.dtype == torch.float32 and out.is_contiguous() and out.numel() >= C
    rc = _lib.lib().faln_channel_sum_nhwc(_lib.ptr(g), _lib.ptr(out), B * H * W, C, Cs, _lib.cur_stream())
    _lib.check(rc, "faln_channel_sum_nhwc")
    return out


def conv3x3_wgrad(g, x, dW, cout=None, cx=None, ci_off=0, stride=1, flags=0, dbias=None):
    """dW[:cout, ci_off:ci_off+cx] += weight gradient of a 3x3 conv (pad 1) on the tcgen05 kernel (csrc/conv_wgrad.cu).
    g: bf16 [B,Cg,Hg,Wg] channels_last (pre-activation gradient), x: bf16 [B,Cxs,H,W] channels_last (the conv's input, or
    one source of a concatenated input), dW: fp32 [Cout,Cin_tot,3,3] in ``torch.channels_last`` memory format (= KRSC in
    memory; typically a view of the flat gradient arena), accumulated in place with split-K fp32 reductions.
    dbias: optional fp32 [>= cout], contiguous: += sum of g over (b, h, w), the layer's bias gradient, computed by the same
    launch from the kernel's spare operand-window slot (no extra pass over g)."""
    g, x = _nhwc(g), _nhwc(x)
    B, Cg, Hg, Wg = g.shape
    _, Cxs, H, W = x.shape
    assert x.shape[0] == B and (Hg, Wg) == ((H - 1) // stride + 1, (W - 1) // stride + 1), (g.shape, x.shape, stride)
    assert dbias is None or (dbias.dtype == torch.float32 and dbias.is_contiguous() and dbias.numel() >= (cout or dW.shape[0]))
    assert dW.dtype == torch.float32 and dW.dim() == 4 and dW.shape[2:] == (3, 3)
    assert dW.permute(0, 2, 3, 1).is_contiguous(), "dW must be channels_last (KRSC memory)"
    cout = cout or dW.shape[0]
    cx = cx or min(Cxs, dW.shape[1] - ci_off)
    assert cout <= dW.shape[0] and ci_off + cx <= dW.shape[1]
    ev = _timed("conv_wgrad", 2 * 9 * cx * cout * B * Hg * Wg, 2 * B * Hg * Wg * Cg + 2 * B * H * W * Cxs + 4 * 9 * cx * cout, 64)
    rc = _lib.lib().faln_conv3x3_wgrad_bias(_lib.ptr(g), _lib.ptr(x), _lib.ptr(dW), _lib.ptr(dbias), B, H, W, Cg, Cxs, cout,
                                            cx, ci_off, dW.shape[1], stride, int(flags), _lib.cur_stream())
    _lib.check(rc, "faln_conv3x3_wgrad_bias")
    if ev is not None:
        ev.record()
    return dW


class _WgradJob(ctypes.Structure):
    _fields_ = ([("g", ctypes.c_void_p), ("x", ctypes.c_void_p), ("dW", ctypes.c_void_p), ("dbias", ctypes.c_void_p)] +
                [(n, ctypes.c_int) for n in ("B", "H", "W", "Cg", "Cxs", "Cout", "Cx", "ci_off", "Cin_tot", "stride")] +
                [("flags", ctypes.c_uint)])


def conv3x3_wgrad_multi(jobs):
    """Several ``conv3x3_wgrad`` calls in as few launches as possible (csrc/conv_wgrad.cu conv3x3_wgrad_multi_kernel: jobs with
    the same kernel configuration share one grid).  jobs: list of dicts with the keyword arguments of ``conv3x3_wgrad``
    (g, x, dW, cout, cx, ci_off, stride, dbias).  Meant for the small-map layers, whose separate launches are latency chains."""
    n = len(jobs)
    if n == 0:
        return
    arr = (_WgradJob * n)()
    keep, flops, nbytes = [], 0, 0
    for i, j in enumerate(jobs):
        g, x, dW = _nhwc(j["g"]), _nhwc(j["x"]), j["dW"]
        stride = j.get("stride", 1)
        B, Cg, Hg, Wg = g.shape
        _, Cxs, H, W = x.shape
        assert x.shape[0] == B and (Hg, Wg) == ((H - 1) // stride + 1, (W - 1) // stride + 1), (g.shape, x.shape, stride)
        assert dW.dtype == torch.float32 and dW.dim() == 4 and dW.shape[2:] == (3, 3) and dW.permute(0, 2, 3, 1).is_contiguous()
        cout = j.get("cout") or dW.shape[0]
        ci_off = j.get("ci_off", 0)
        cx = j.get("cx") or min(Cxs, dW.shape[1] - ci_off)
        dbias = j.get("dbias")
        assert cout <= dW.shape[0] and ci_off + cx <= dW.shape[1]
        assert dbias is None or (dbias.dtype == torch.float32 and dbias.is_contiguous() and dbias.numel() >= cout)
        keep.extend((g, x))
        a = arr[i]
        a.g, a.x, a.dW = g.data_ptr(), x.data_ptr(), dW.data_ptr()
        a.dbias = dbias.data_ptr() if dbias is not None else None
        a.B, a.H, a.W, a.Cg, a.Cxs, a.Cout, a.Cx, a.ci_off, a.Cin_tot, a.stride = B, H, W, Cg, Cxs, cout, cx, ci_off, dW.shape[1], stride
        a.flags = 0
        flops += 2 * 9 * cx * cout * B * Hg * Wg
        nbytes += 2 * B * Hg * Wg * Cg + 2 * B * H * W * Cxs + 4 * 9 * cx * cout
    if not keep[0].is_cuda:
        raise RuntimeError("fal_net_b200 kernels need CUDA tensors (no CPU fallback)")
    ev = _timed("conv_wgrad", flops, nbytes, 64)
    rc = _lib.lib().faln_conv3x3_wgrad_multi(ctypes.cast(arr, ctypes.c_void_p), n, _lib.cur_stream())
    _lib.check(rc, "faln_conv3x3_wgrad_multi")
    if ev is not None:
        ev.record()


def conv3x3_wgrad_up2_multi(jobs):
    """Several ``conv3x3_wgrad_up2`` calls as one grid (csrc/conv_wgrad.cu conv3x3_wgrad_up2_multi_kernel); jobs: list of dicts
    with keys g, x, dW, cout (optional)."""
    n = len(jobs)
    if n == 0:
        return
    if n == 1:
        j = jobs[0]
        conv3x3_wgrad_up2(j["g"], j["x"], j["dW"], cout=j.get("cout"))
        return
    arr = (_WgradJob * n)()
    keep, flops, nbytes = [], 0, 0
    for i, j in enumerate(jobs):
        g, x, dW = _nhwc(j["g"]), _nhwc(j["x"]), j["dW"]
        B, Cg, Hg, Wg = g.shape
        _, Cxs, H, W = x.shape
        assert x.shape[0] == B and (Hg, Wg) == (2 * H, 2 * W) and Cg % 64 == 0 and Cxs % 64 == 0, (g.shape, x.shape)
        assert dW.dtype == torch.float32 and dW.dim() == 4 and dW.shape[2:] == (3, 3) and dW.permute(0, 2, 3, 1).is_contiguous()
        assert g.is_cuda
        cout = j.get("cout") or dW.shape[0]
        cx = min(Cxs, dW.shape[1])
        keep.extend((g, x))
        a = arr[i]
        a.g, a.x, a.dW, a.dbias = g.data_ptr(), x.data_ptr(), dW.data_ptr(), None
        a.B, a.H, a.W, a.Cg, a.Cxs, a.Cout, a.Cx, a.ci_off, a.Cin_tot, a.stride = B, H, W, Cg, Cxs, cout, cx, 0, dW.shape[1], 1
        a.flags = 0
        flops += 2 * 9 * cx * cout * B * Hg * Wg
        nbytes += 2 * B * Hg * Wg * Cg + 2 * B * H * W * Cxs + 4 * 9 * cx * cout
    ev = _timed("conv_wgrad", flops, nbytes, 64)
    rc = _lib.lib().faln_conv3x3_wgrad_up2_multi(ctypes.cast(arr, ctypes.c_void_p), n, _lib.cur_stream())
    _lib.check(rc, "faln_conv3x3_wgrad_up2_multi")
    if ev is not None:
        ev.record()


def conv3x3_wgrad_up2(g, x, dW, cout=None, cx=None, ci_off=0):
    """dW[:cout, ci_off:ci_off+cx] += weight gradient of "nearest 2x up-sampling, then conv3x3" (the reference's deconv block,
    /root/reference/models/FAL_netB.py:51-60) taken from the LOW-resolution input: g bf16 [B,Cg,2H,2W] channels_last
    (pre-activation gradient on the up-sampled grid), x bf16 [B,Cxs,H,W] channels_last, dW as in ``conv3x3_wgrad``.
    Cg and Cxs must be multiples of 64 (csrc/conv_wgrad.cu: conv3x3_wgrad_up2_kernel)."""
    g, x = _nhwc(g), _nhwc(x)
    B, Cg, Hg, Wg = g.shape
    _, Cxs, H, W = x.shape
    assert x.shape[0] == B and (Hg, Wg) == (2 * H, 2 * W), (g.shape, x.shape)
    assert Cg % 64 == 0 and Cxs % 64 == 0, (Cg, Cxs)
    assert dW.dtype == torch.float32 and dW.dim() == 4 and dW.shape[2:] == (3, 3)
    assert dW.permute(0, 2, 3, 1).is_contiguous(), "dW must be channels_last (KRSC memory)"
    cout = cout or dW.shape[0]
    cx = cx or min(Cxs, dW.shape[1] - ci_off)
    assert cout <= dW.shape[0] and ci_off + cx <= dW.shape[1]
    # reference formulation of the work (nine taps on the up-sampled grid), like the folded forward / data gradient
    ev = _timed("conv_wgrad", 2 * 9 * cx * cout * B * Hg * Wg, 2 * B * Hg * Wg * Cg + 2 * B * H * W * Cxs + 4 * 9 * cx * cout, 64)
    rc = _lib.lib().faln_conv3x3_wgrad_up2(_lib.ptr(g), _lib.ptr(x), _lib.ptr(dW), B, H, W, Cg, Cxs, cout, cx, ci_off,
                                           dW.shape[1], _lib.cur_stream())
    _lib.check(rc, "faln_conv3x3_wgrad_up2")
    if ev is not None:
        ev.record()
    return dW


def border_sums(g, C=None):
    """[B,3,3,C] fp32 sums of the bf16 channels_last map g over the 3x3 border classes (first / interior / last row x col)."""
    g = _nhwc(g)
    B, Cs, H, W = g.shape
    C = C or Cs
    out = torch.zeros(B, 3, 3, C, device=g.device, dtype=torch.float32)
    rc = _lib.lib().faln_border_sum_nhwc(_lib.ptr(g), _lib.ptr(out), B, H, W, C, Cs, _lib.cur_stream())
    _lib.check(rc, "faln_border_sum_nhwc")
    return out


_TAP_VALID: dict = {}


def _tap_validity(device, last_clipped: bool) -> torch.Tensor:
    """[3 taps, 3 border classes] 0/1: tap 0 of the first row/col reads index -1; tap 2 of the last row/col reads index
    n_in when ``last_clipped``.  Cached on the device (no host-to-device copy inside a CUDA-graph capture)."""
    key = (str(device), bool(last_clipped))
    m = _TAP_VALID.get(key)
    if m is None:
        m = torch.ones(3, 3, device=device)        # built with device-side fills only: legal inside a graph capture
        m[0, 0].fill_(0.0)
        if last_clipped:
            m[2, 2].fill_(0.0)
        _TAP_VALID[key] = m
    return m


def const_channel_wgrad(g, value, in_hw, stride, cout):
    """Weight gradient [cout,3,3] of a spatially constant input channel with per-sample value ``value`` [B]: for each tap
    the sum of g over the output pixels whose tap lands inside the image -- a combination of the nine border-class sums."""
    B, Cs, Hg, Wg = g.shape
    H, W = in_hw
    S = border_sums(g, cout)                                               # [B,3,3,C]
    S = torch.einsum("b,brcn->rcn", value.float(), S)                      # [3(row class),3(col class),C]

    vh = _tap_validity(g.device, stride * (Hg - 1) + 1 > H - 1)
    vw = _tap_validity(g.device, stride * (Wg - 1) + 1 > W - 1)
    return torch.einsum("hr,wc,rcn->nhw", vh, vw, S)
