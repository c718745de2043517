"""VGG19 perceptual slices (/root/reference/loss_functions.py:7-44) on the tcgen05 convolution kernels.

Activations are bf16 tensors of logical shape [B,C,H,W] in ``torch.channels_last`` memory format, i.e. NHWC in memory --
the layout the implicit-GEMM kernels (csrc/conv_tc.cu) consume through TMA.

The three slices are ONE autograd node with a hand-scheduled backward (like fal_net_b200.backbone): the network is
frozen (reference :33-34), so backward is the data-gradient chain only -- ``faln_conv3x3_dgrad`` with the ReLU
derivative, the pool-output gradients of the shallower slices and the max-pool routing fused into epilogues / one small
kernel each -- down to the fp32 gradient of the input image.  No cuDNN, no ATen elementwise glue.  Nothing here runs
without CUDA.
"""
from __future__ import annotations

import torch

from . import conv_native as CN
from . import layout

# Re-packed bf16 weights are cached ON the parameter object, tagged with (param._version, generation): parameters that do
# not change (frozen Stage-1 model, VGG, inference) are packed once; the trainer bumps GENERATION after each fused-Adam
# step because that kernel updates the arena behind torch's version counter.  (A dict keyed by id(param) is wrong: ids
# are recycled when a model is freed.)
GENERATION = [0]
LIBRARY_CALLS = {"conv_backward": 0}     # stays 0: kept so bench.py can report that no library convolution ran


def invalidate_packed_weights():
    GENERATION[0] += 1


CL = torch.channels_last
_ACT = CN.ACT

# ------------------------------------------------------------------------------------------------
# VGG19 features[0:19] (the three pooled activations of /root/reference/loss_functions.py:21-29,36-44)
# ------------------------------------------------------------------------------------------------
VGG_CFG = (64, 64, "M", 128, 128, "M", 256, 256, 256, 256, "M")


def _vgg_pack(w, tag, fn):
    key = (w._version, w.data_ptr(), tag)
    cache = getattr(w, "_faln_packs", None)
    if cache is None:
        cache = w._faln_packs = {}
    if cache.get("key_" + tag) != key:
        cache["key_" + tag] = key
        cache[tag] = fn()
    return cache[tag]


def _vgg_forward(ws, x, tape=None):
    outs, i = [], 0
    for v in VGG_CFG:
        if v == "M":
            y = CN.maxpool2(x)
            if tape is not None:
                tape.append(("pool", x))
            x = y
            outs.append(x)
        else:
            w, b = ws[i]
            if i == 0:
                y = CN.stem_conv(x, w, b, _ACT["relu"])
            else:
                y = CN.conv3x3_fwd(x, _vgg_pack(w, "fwd", lambda: CN.pack_weight(w)), b, 1, _ACT["relu"])
            if tape is not None:
                tape.append(("conv", i, y))
            x = y
            i += 1
    return outs


class _VggFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, *flat_ws):
        ws = list(zip(flat_ws[0::2], flat_ws[1::2]))
        tape = []
        outs = _vgg_forward(ws, x, tape)
        ctx.tape, ctx.ws, ctx.in_shape = tape, ws, x.shape
        return tuple(outs)

    @staticmethod
    def backward(ctx, *g_outs):
        tape, ws = ctx.tape, ctx.ws
        ctx.tape = None
        g_outs = [None if t is None else t.to(torch.bfloat16).contiguous(memory_format=CL) for t in g_outs]
        g = None                                     # gradient w.r.t. the tensor produced by the tape entry being visited
        for k in range(len(tape) - 1, -1, -1):
            ent = tape[k]
            if ent[0] == "pool":
                x = ent[1]                           # the pool's input = ReLU output of the previous conv
                if g is None:                        # deepest slice: only its own output gradient arrives here
                    g = g_outs.pop()
                    if g is None:
                        g = torch.zeros((x.shape[0], x.shape[1], x.shape[2] // 2, x.shape[3] // 2), device=x.device,
                                        dtype=torch.bfloat16).contiguous(memory_format=CL)
                # max-pool routing fused with ReLU'(x): g becomes the gradient w.r.t. the previous conv's PRE-activation
                g = CN.maxpool2_bwd(x, g, dact=2)
            else:
                _, i, y = ent
                hw = (y.shape[2], y.shape[3])
                if i == 0:
                    # stem (3 -> 64): data gradient on the tensor-core path with the 3 input channels padded to 32
                    w = ws[0][0]
                    wd = _vgg_pack(w, "dgrad", lambda: CN.pack_weight_dgrad(w, cin_pad=32))
                    gx = CN.conv3x3_dgrad(g, wd, hw)
                    gx = layout.nhwc_bf16_to_planar(gx.permute(0, 2, 3, 1), 3, pitch=hw[1])
                    return (gx.contiguous(),) + (None,) * (2 * len(ws))
                w = ws[i][0]
                wd = _vgg_pack(w, "dgrad", lambda: CN.pack_weight_dgrad(w))
                prev = tape[k - 1]
                if prev[0] == "conv":                # producer is conv + ReLU: fuse ReLU'(its saved output)
                    g = CN.conv3x3_dgrad(g, wd, hw, dact=2, ysave=prev[2])
                else:                                # producer is a pool, whose output is also a slice output: the
                    g_own = g_outs.pop() if g_outs else None               # epilogue adds that slice's own gradient
                    g = CN.conv3x3_dgrad(g, wd, hw, residual=g_own)
        raise AssertionError("unreachable")


def vgg_features(ws, x):
    """ws: list of (weight fp32 [Co,Ci,3,3], bias fp32 [Co]) for the 8 convs; x fp32 NCHW image [B,3,H,W].
    Returns the three pooled activations (bf16 channels_last)."""
    if not x.is_cuda:
        raise RuntimeError("fal_net_b200 VGG features run on CUDA (sm_100a) only; there is no CPU path")
    if torch.is_grad_enabled() and x.requires_grad:
        flat = [t for wb in ws for t in wb]
        return list(_VggFn.apply(x, *flat))
    return _vgg_forward(ws, x)
