"""CPU oracle for the FAL-net hot path  --  TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference``
legs of ``bench.py`` may import this file.  Nothing under ``fal_net_b200/`` imports it.

It restates, with stock ``torch`` CPU ops in fp32 (the reference's own arithmetic lives in
PyTorch: ``conv2d``, ``grid_sample``, ``affine_grid``, ``softmax``, ``interpolate``; see
SURVEY.md 8c), the algorithm of

  * ``FAL_netB.forward``            -> /root/reference/models/FAL_netB.py:200-297
  * ``BackBone.forward``            -> /root/reference/models/FAL_netB.py:140-176
  * the losses                      -> /root/reference/loss_functions.py:52-109
  * the VGG19 perceptual slices     -> /root/reference/loss_functions.py:7-44
  * the Stage-1 / Stage-2 / Test step bodies
                                    -> /root/reference/Train_Stage1_K.py:223-262,
                                       /root/reference/Train_Stage2_K.py:233-331,
                                       /root/reference/Test_KITTI.py:163-208,287-300

Two flavours of the MED section are given:

  ``med_forward_ops``     replays the reference's op sequence (``affine_grid`` + N x
                          ``grid_sample`` + ``softmax``) and is BIT-IDENTICAL to the reference on
                          CPU (pinned by tests/golden/make_golden.py, run in the build container
                          against the imported reference).
  ``med_forward_closed``  the closed-form gather formulation the CUDA kernels implement
                          (SURVEY.md A.2), including the reference's fp32 normalised-coordinate
                          rounding; ``med_backward_closed`` is its analytic adjoint (A.3).

Parity status: the reference ships no tests / golden vectors of its own, so this
oracle is pinned against OUTPUTS OF THE REFERENCE ITSELF generated in the build container by
``tests/golden/make_golden.py`` and committed under ``tests/golden/*.npz``.
"""
from __future__ import annotations

import math
from collections import OrderedDict

import numpy as np
import torch
import torch.nn.functional as F

# ----------------------------------------------------------------------------------------------
# Parameters
# ----------------------------------------------------------------------------------------------

# (name, cin, cout, stride, bias) in construction order of BackBone.__init__
# (/root/reference/models/FAL_netB.py:99-128); residual blocks own two bias-free convs.
ENCODER = [("conv0", 3, 32, 1), ("conv1", 33, 64, 2), ("conv2", 64, 128, 2), ("conv3", 128, 256, 2),
           ("conv4", 256, 256, 2), ("conv5", 256, 256, 2), ("conv6", 256, 512, 2)]
DECODER = [(6, 512, 256, 256 + 256, 256), (5, 256, 128, 128 + 256, 256), (4, 256, 128, 128 + 256, 256),
           (3, 256, 128, 128 + 128, 128), (2, 128, 64, 64 + 64, 64), (1, 64, 64, 32 + 64, None)]


# Variants.  FAL_netC (/root/reference/models/FAL_netC.py:110-120,185): wider bottleneck, encoder-decoder
# registered as ``synth``.  FAL_netA (/root/reference/models/FAL_netA.py:73-76,99-126,183,264): narrower, 3x1 / 1x3 residual
# kernels, no amask_conv, registered as ``BackBone``, maskR sampled with grid_sample's default align_corners=False.
VARIANTS = {
    "FAL_netB": dict(prefix="backbone", enc=ENCODER, dec=DECODER, separable=False, amask=True, maskr_align=True),
    "FAL_netC": dict(prefix="synth",
                     enc=[("conv0", 3, 32, 1), ("conv1", 33, 64, 2), ("conv2", 64, 128, 2), ("conv3", 128, 256, 2),
                          ("conv4", 256, 256, 2), ("conv5", 256, 512, 2), ("conv6", 512, 512, 2)],
                     dec=[(6, 512, 256, 256 + 512, 512), (5, 512, 256, 256 + 256, 256), (4, 256, 128, 128 + 256, 256),
                          (3, 256, 128, 128 + 128, 128), (2, 128, 64, 64 + 64, 64), (1, 64, 64, 32 + 64, None)],
                     separable=False, amask=True, maskr_align=True),
    "FAL_netA": dict(prefix="BackBone",
                     enc=[("conv0", 3, 32, 1), ("conv1", 33, 64, 2), ("conv2", 64, 128, 2), ("conv3", 128, 128, 2),
                          ("conv4", 128, 256, 2), ("conv5", 256, 256, 2), ("conv6", 256, 256, 2)],
                     dec=[(6, 256, 128, 128 + 256, 256), (5, 256, 128, 128 + 256, 256), (4, 256, 128, 128 + 128, 128),
                          (3, 128, 64, 128 + 64, 128), (2, 128, 64, 64 + 64, 64), (1, 64, 64, 32 + 64, None)],
                     separable=True, amask=False, maskr_align=False),
}


def param_shapes(no_levels: int = 49, variant: str = "FAL_netB") -> "OrderedDict[str, tuple]":
    """state_dict keys and shapes in registration order."""
    v = VARIANTS[variant]
    pf = v["prefix"]
    k1, k2 = ((3, 1), (1, 3)) if v["separable"] else ((3, 3), (3, 3))
    sh = OrderedDict()
    for name, cin, cout, _ in v["enc"]:
        sh[f"{pf}.{name}.0.weight"] = (cout, cin, 3, 3)
        sh[f"{pf}.{name}.0.bias"] = (cout,)
        sh[f"{pf}.{name}_1.conv1.weight"] = (cout, cout) + k1
        sh[f"{pf}.{name}_1.conv2.weight"] = (cout, cout) + k2
    for lvl, din, dout, iin, iout in v["dec"]:
        sh[f"{pf}.deconv{lvl}.conv1.weight"] = (dout, din, 3, 3)
        if iout is not None:
            sh[f"{pf}.iconv{lvl}.0.weight"] = (iout, iin, 3, 3)
            sh[f"{pf}.iconv{lvl}.0.bias"] = (iout,)
        else:
            sh[f"{pf}.iconv1.weight"] = (no_levels, iin, 3, 3)
    if v["amask"]:
        sh[f"{pf}.amask_conv.0.weight"] = (48, 96, 3, 3)
        sh[f"{pf}.amask_conv.0.bias"] = (48,)
        sh[f"{pf}.amask_conv.2.weight"] = (1, 48, 3, 3)
    sh["conv0.weight"] = (no_levels, no_levels, 1, 1)
    sh["conv0.bias"] = (no_levels,)
    return sh


def variant_of(p) -> str:
    """Which variant a parameter dict belongs to (from its key prefix)."""
    k = next(iter(p.keys()))
    return {"backbone": "FAL_netB", "synth": "FAL_netC", "BackBone": "FAL_netA"}[k.split(".")[0]]


def init_params(no_levels: int = 49, seed: int = 0) -> "OrderedDict[str, torch.Tensor]":
    """Kaiming-normal (fan_in, gain sqrt(2)) weights, zero biases
    (/root/reference/models/FAL_netB.py:130-138,191-192).  NOT the same random stream as the
    reference constructor (that one is reproduced by the product module and pinned by golden
    checksums); the oracle only needs *a* valid parameter set."""
    g = torch.Generator().manual_seed(seed)
    out = OrderedDict()
    for k, shp in param_shapes(no_levels).items():
        if k.endswith("bias"):
            out[k] = torch.zeros(shp)
        else:
            fan_in = shp[1] * shp[2] * shp[3]
            out[k] = torch.randn(shp, generator=g) * math.sqrt(2.0 / fan_in)
    return out


# ----------------------------------------------------------------------------------------------
# Backbone  (/root/reference/models/FAL_netB.py:140-176)
# ----------------------------------------------------------------------------------------------

def _pad_of(w):
    return ((w.shape[2] - 1) // 2, (w.shape[3] - 1) // 2)


def _res_block(p, prefix, x):
    # /root/reference/models/FAL_netB.py:78-80 (3x3 kernels); FAL_netA.py:73-80 (3x1 then 1x3, "same" padding)
    w1, w2 = p[prefix + ".conv1.weight"], p[prefix + ".conv2.weight"]
    y = F.elu(F.conv2d(x, w1, None, 1, _pad_of(w1)))
    y = F.conv2d(y, w2, None, 1, _pad_of(w2))
    return F.elu(y + x)


def backbone_forward(p, x, flow, collect=None):
    """x [B,3,H,W], flow [B,1,H,W] -> dlog [B,N,H,W]."""
    v = VARIANTS[variant_of(p)]
    pf = v["prefix"]
    skips = []
    h = x
    for i, (name, _, _, stride) in enumerate(v["enc"]):
        if i == 1:
            h = torch.cat((h, flow), 1)                                       # :145
        h = F.elu(F.conv2d(h, p[f"{pf}.{name}.0.weight"], p[f"{pf}.{name}.0.bias"], stride, 1))
        h = _res_block(p, f"{pf}.{name}_1", h)
        skips.append(h)
        if collect is not None:
            collect[name] = h
    h = skips[6]
    for lvl, _, _, _, iout in v["dec"]:
        skip = skips[lvl - 1]
        u = F.interpolate(h, size=(skip.shape[2], skip.shape[3]), mode="nearest")   # :58
        u = F.elu(F.conv2d(u, p[f"{pf}.deconv{lvl}.conv1.weight"], None, 1, 1))      # :59
        c = torch.cat((u, skip), 1)
        if iout is not None:
            h = F.elu(F.conv2d(c, p[f"{pf}.iconv{lvl}.0.weight"], p[f"{pf}.iconv{lvl}.0.bias"], 1, 1))
        else:
            h = F.conv2d(c, p[f"{pf}.iconv1.weight"], None, 1, 1)               # :127,174 no activation
        if collect is not None:
            collect[f"iconv{lvl}"] = h
    return h


# ----------------------------------------------------------------------------------------------
# MED section
# ----------------------------------------------------------------------------------------------

def level_tables(min_disp, max_disp, no_levels, W):
    """Per-sample level tables, with the reference's fp32 expressions
    (/root/reference/models/FAL_netB.py:204-205,224-225,241).
    min_disp/max_disp: [B,1,1] fp32.  Returns d [B,N] (disparity of level n, pixels) and
    x_of [B,N] (normalised-grid x offset of level n)."""
    x_pix_min = 2 * min_disp / W
    x_pix_max = 2 * max_disp / W
    d, xo = [], []
    for n in range(no_levels):
        c = n / (no_levels - 1)
        d.append(max_disp * torch.exp(torch.log(max_disp / min_disp) * (c - 1)))
        xo.append(x_pix_max * torch.exp(torch.log(x_pix_max / x_pix_min) * (c - 1)))
    return torch.cat(d, 2).squeeze(1), torch.cat(xo, 2).squeeze(1)


def identity_grid(B, C, H, W, device=None, align_corners=True):
    th = torch.zeros(B, 2, 3, device=device)
    th[:, 0, 0] = 1
    th[:, 1, 1] = 1
    return F.affine_grid(th, [B, C, H, W], align_corners=align_corners)


def med_forward_ops(dlog0, image, min_disp, max_disp, ret_disp=True, ret_subocc=False, ret_pan=False,
                    maskr_align=True):
    """The reference's op sequence on (dlog0, image): /root/reference/models/FAL_netB.py:216-297.
    Returns the same thing as FAL_net.forward (a tensor when only ret_disp, else a list ordered
    [pan?, disp?, maskL?, maskR?])."""
    B, C, H, W = image.shape
    N = dlog0.shape[1]
    x_pix_min = 2 * min_disp / W
    x_pix_max = 2 * max_disp / W
    sm0 = torch.softmax(dlog0, dim=1)
    disp = None
    if ret_disp:
        disp = 0
        for n in range(N):
            c = n / (N - 1)
            w = max_disp * torch.exp(torch.log(max_disp / min_disp) * (c - 1))
            disp = disp + w.unsqueeze(1) * sm0[:, n].unsqueeze(1)
    if ret_disp and not ret_subocc and not ret_pan:
        return disp
    grid = identity_grid(B, C, H, W, image.device)

    def shifted(n, sign):
        c = n / (N - 1)
        x_of = x_pix_max * torch.exp(torch.log(x_pix_max / x_pix_min) * (c - 1))
        g = grid.clone()
        if sign > 0:
            g[..., 0] = g[..., 0] + x_of                                         # :242-243
        else:
            g[..., 0] = g[..., 0] - x_of                                         # :270-271
        return g

    planes = [F.grid_sample(dlog0[:, n].unsqueeze(1), shifted(n, +1), align_corners=True) for n in range(N)]
    Dprob = torch.softmax(torch.cat(planes, 1), dim=1)                           # :248
    pan, maskR, maskL = 0, 0, 0
    for n in range(N):
        g = shifted(n, +1)
        if ret_subocc:
            with torch.no_grad():
                # FAL_netA.py:264 calls grid_sample WITHOUT align_corners for maskR (default False); B / C pass True
                maskR = maskR + F.grid_sample(sm0[:, n].unsqueeze(1).detach(), g, align_corners=bool(maskr_align))
                maskL = maskL + F.grid_sample(Dprob[:, n].unsqueeze(1).detach(), shifted(n, -1),
                                              align_corners=True)
        if ret_pan:
            pan = pan + F.grid_sample(image, g, align_corners=True) * Dprob[:, n].unsqueeze(1)
    out = []
    if ret_pan:
        out.append(pan)
    if ret_disp:
        out.append(disp)
    if ret_subocc:
        out.append(torch.clamp(maskL, max=1.0))
        out.append(torch.clamp(maskR, max=1.0))
    return out


def _coords(g0x, x_of, W, sign):
    """fp32 replay of the sampling coordinate: normalised grid value + offset,
    then ATen's un-normalisation ((g + 1) / 2) * (W - 1) in that op order.
    g0x [W], x_of [B,N] -> ix [B,N,W] fp32, x0 [B,N,W] int64, a [B,N,W] fp32."""
    if sign > 0:
        gx = g0x.view(1, 1, -1) + x_of.unsqueeze(-1)
    else:
        gx = g0x.view(1, 1, -1) - x_of.unsqueeze(-1)
    ix = ((gx + 1.0) * 0.5) * float(W - 1)
    x0f = torch.floor(ix)
    return ix, x0f.long(), ix - x0f


def _gather_shift(f, x0, a):
    """f [B,N,H,W] (or [B,1,H,W]); x0,a [B,N,W].  Zero-padded two-tap horizontal resample."""
    B, N, W = x0.shape
    H = f.shape[2]
    f = f.expand(B, N, H, f.shape[3])
    i0 = x0.unsqueeze(2).expand(B, N, H, W)
    i1 = i0 + 1
    v0 = ((i0 >= 0) & (i0 <= W - 1)).to(f.dtype)
    v1 = ((i1 >= 0) & (i1 <= W - 1)).to(f.dtype)
    t0 = torch.gather(f, 3, i0.clamp(0, W - 1)) * v0
    t1 = torch.gather(f, 3, i1.clamp(0, W - 1)) * v1
    aa = a.unsqueeze(2)
    return t0 * (1 - aa) + t1 * aa


def med_forward_closed(dlog0, image, d, x_of, want_masks=True):
    """Closed-form MED forward with the fp32 coordinate replay.
    dlog0 [B,N,H,W], image [B,3,H,W], d/x_of [B,N] from ``level_tables``.
    Returns dict(pan, disp, maskL, maskR, lse0, lsew)."""
    B, N, H, W = dlog0.shape
    g0x = identity_grid(1, 1, 1, W)[0, 0, :, 0].to(dlog0.dtype)
    _, x0p, ap = _coords(g0x, x_of, W, +1)
    p0 = torch.softmax(dlog0, 1)
    disp = (p0 * d.view(B, N, 1, 1)).sum(1, keepdim=True)
    wl = _gather_shift(dlog0, x0p, ap)
    P = torch.softmax(wl, 1)
    pan = torch.zeros_like(image)
    for c in range(image.shape[1]):
        pan[:, c] = (_gather_shift(image[:, c:c + 1], x0p, ap) * P).sum(1)
    out = dict(pan=pan, disp=disp, lse0=torch.logsumexp(dlog0, 1, keepdim=True),
               lsew=torch.logsumexp(wl, 1, keepdim=True))
    if want_masks:
        _, x0m, am = _coords(g0x, x_of, W, -1)
        out["maskR"] = torch.clamp(_gather_shift(p0, x0p, ap).sum(1, keepdim=True), max=1.0)
        out["maskL"] = torch.clamp(_gather_shift(P, x0m, am).sum(1, keepdim=True), max=1.0)
    return out


def med_backward_closed(dlog0, image, d, x_of, g_pan, g_disp):
    """Analytic adjoint of ``med_forward_closed`` w.r.t. dlog0; only pan and disp
    carry gradient.  Scatter form (index_add) -- the CUDA kernel computes the same sums as a
    windowed gather."""
    B, N, H, W = dlog0.shape
    g0x = identity_grid(1, 1, 1, W)[0, 0, :, 0].to(dlog0.dtype)
    _, x0, a = _coords(g0x, x_of, W, +1)
    p0 = torch.softmax(dlog0, 1)
    disp = (p0 * d.view(B, N, 1, 1)).sum(1, keepdim=True)
    wl = _gather_shift(dlog0, x0, a)
    P = torch.softmax(wl, 1)
    dP = torch.zeros_like(P)
    for c in range(image.shape[1]):
        dP = dP + g_pan[:, c:c + 1] * _gather_shift(image[:, c:c + 1], x0, a)
    dot = (P * dP).sum(1, keepdim=True)
    dwl = P * (dP - dot)
    gL = torch.zeros(B, N, H, W + 2, dtype=dlog0.dtype)        # one guard column either side
    i0 = (x0.unsqueeze(2).expand(B, N, H, W) + 1)
    aa = a.unsqueeze(2)
    v0 = ((i0 - 1 >= 0) & (i0 - 1 <= W - 1)).to(dlog0.dtype)
    v1 = ((i0 >= 0) & (i0 <= W - 1)).to(dlog0.dtype)
    gL.scatter_add_(3, i0.clamp(0, W + 1), dwl * (1 - aa) * v0)
    gL.scatter_add_(3, (i0 + 1).clamp(0, W + 1), dwl * aa * v1)
    gL = gL[..., 1:W + 1]
    gL = gL + p0 * g_disp * (d.view(B, N, 1, 1) - disp)
    return gL


def falnet_forward(p, image, min_disp, max_disp, ret_disp=True, ret_subocc=False, ret_pan=False):
    """FAL_net.forward (/root/reference/models/FAL_netB.py:200-297) on a parameter dict."""
    B, C, H, W = image.shape
    flow = torch.ones(B, 1, H, W, dtype=image.dtype, device=image.device)
    flow[:, 0] = max_disp * flow[:, 0] / 100                                     # :208-209
    dlog = backbone_forward(p, image, flow)
    dlog0 = F.conv2d(dlog, p["conv0.weight"], p["conv0.bias"])                   # :215
    return med_forward_ops(dlog0, image, min_disp, max_disp, ret_disp, ret_subocc, ret_pan,
                           maskr_align=VARIANTS[variant_of(p)]["maskr_align"])


# ----------------------------------------------------------------------------------------------
# Losses  (/root/reference/loss_functions.py)
# ----------------------------------------------------------------------------------------------

VGG_CFG = [64, 64, "M", 128, 128, "M", 256, 256, 256, 256, "M"]       # torchvision vgg19 features[0:19]
VGG_SLICE_END = {1: 0, 4: 1, 9: 2}                                    # conv index after which a slice ends (its pool)


def vgg_shapes():
    sh, cin = [], 3
    for v in VGG_CFG:
        if v != "M":
            sh.append((v, cin, 3, 3))
            cin = v
    return sh


def init_vgg(seed: int = 2):
    """Seeded random stand-in for the ImageNet VGG19 weights (unobtainable offline).
    Kaiming-normal fan_out like torchvision's own initialiser, zero bias."""
    g = torch.Generator().manual_seed(seed)
    ws = []
    for shp in vgg_shapes():
        fan_out = shp[0] * 9
        ws.append((torch.randn(shp, generator=g) * math.sqrt(2.0 / fan_out), torch.zeros(shp[0])))
    return ws


def vgg_features(ws, x):
    """The three activations the reference's Vgg19_pc returns (each slice ends in its max-pool):
    /root/reference/loss_functions.py:21-29,36-44."""
    outs, i = [], 0
    for v in VGG_CFG:
        if v == "M":
            x = F.max_pool2d(x, 2, 2)
            outs.append(x)
        else:
            x = F.relu(F.conv2d(x, ws[i][0], ws[i][1], 1, 1))
            i += 1
    return outs


def perceptual_loss(a, b):
    # /root/reference/loss_functions.py:59-67
    return sum(torch.mean((x - y) ** 2) for x, y in zip(a, b))


def rec_loss(mask, synth, label, vgg_label, a_p, vgg_ws=None):
    # /root/reference/loss_functions.py:52-56
    loss = torch.mean(mask * torch.abs(synth - label))
    if a_p > 0 and vgg_label is not None:
        loss = loss + a_p * perceptual_loss(vgg_features(vgg_ws, mask * synth + (1 - mask) * label), vgg_label)
    return loss


_RGB_MEAN = (0.411, 0.432, 0.45)


def smoothness(img, disp, gamma=1.0):
    """Edge-aware smoothness, zero-padded stencils (/root/reference/loss_functions.py:70-101)."""
    mean = torch.tensor(_RGB_MEAN, dtype=img.dtype).view(1, 3, 1, 1)
    x = img + mean
    gray = (0.299 * x[:, 0] + 0.587 * x[:, 1] + 0.114 * x[:, 2]).unsqueeze(1)          # :104-109

    def stencil(t, rows):
        k = torch.tensor(rows, dtype=t.dtype).view(1, 1, 3, 3)
        return F.conv2d(t, k, padding=1)

    dx_img = stencil(gray, [[0, 0, 0], [-1, 2, -1], [0, 0, 0]])
    dy_img = stencil(gray, [[0, -1, 0], [0, 2, 0], [0, -1, 0]])
    dx_d = stencil(disp, [[0, 0, 0], [0, 1, -1], [0, 0, 0]])
    dy_d = stencil(disp, [[0, -1, 0], [0, 1, 0], [0, 0, 0]])
    dx1_d = stencil(disp, [[0, 0, 0], [-1, 1, 0], [0, 0, 0]])
    dy1_d = stencil(disp, [[0, 0, 0], [0, 1, 0], [0, -1, 0]])
    return torch.mean((dx_d.abs() + dx1_d.abs()) * torch.exp(-gamma * dx_img.abs()) +
                      (dy_d.abs() + dy1_d.abs()) * torch.exp(-gamma * dy_img.abs()))


# ----------------------------------------------------------------------------------------------
# Step bodies
# ----------------------------------------------------------------------------------------------

def stage1_loss(p, left, right, min_disp, max_disp, a_p=0.0, a_sm=0.2 * 2 / 512, vgg_ws=None):
    """/root/reference/Train_Stage1_K.py:236-258.  Returns (loss, rec_loss, sm_loss, pan, disp)."""
    W = left.shape[3]
    pan, disp = falnet_forward(p, left, min_disp, max_disp, ret_disp=True, ret_pan=True)
    vgg_right = vgg_features(vgg_ws, right) if a_p > 0 else None
    rec = rec_loss(1, pan, right, vgg_right, a_p, vgg_ws)
    c0 = int(0.20 * W)
    sm = smoothness(left[..., c0:], disp[..., c0:], gamma=2) if a_sm > 0 else 0
    return rec + a_sm * sm, rec, sm, pan, disp


def grid_flip(x, align_corners=True):
    """Horizontal flip the way the reference does it: bilinear grid_sample on a negated identity
    grid (/root/reference/Train_Stage2_K.py:247-253).  Not bit-exact w.r.t. torch.flip
    (SURVEY.md Appendix B)."""
    B, C, H, W = x.shape
    g = identity_grid(B, C, H, W, x.device, align_corners).clone()
    g[..., 0] = -g[..., 0]
    return F.grid_sample(x, g, align_corners=align_corners)


def stage2_loss(p, p_fix, left, right, min_disp, max_disp, a_p=0.01, a_sm=0.4 * 2 / 512, a_mr=1.0,
                vgg_ws=None, flip=grid_flip):
    """/root/reference/Train_Stage2_K.py:247-327.  ``flip`` is injectable so tests can hand both
    sides an exact index flip."""
    B, C, H, W = left.shape
    mn2, mx2 = torch.cat((min_disp, min_disp), 0), torch.cat((max_disp, max_disp), 0)
    if a_mr > 0:
        with torch.no_grad():
            dfix = falnet_forward(p_fix, torch.cat((flip(left), right), 0), mn2, mx2)
            mldisp = flip(dfix[:B]).detach()
            mrdisp = dfix[B:].detach()
    pan, disp, mask0, mask1 = falnet_forward(p, torch.cat((left, flip(right)), 0), mn2, mx2,
                                             ret_disp=True, ret_pan=True, ret_subocc=True)
    rpan, lpan = pan[:B], flip(pan[B:])
    ldisp, rdisp = disp[:B], flip(disp[B:])
    lmask, rmask = mask0[:B], flip(mask0[B:])
    rlmask, lrmask = mask1[:B], flip(mask1[B:])
    vgg_right = vgg_features(vgg_ws, right) if a_p > 0 else None
    vgg_left = vgg_features(vgg_ws, left) if a_p > 0 else None
    c20, c80 = int(0.20 * W), int(0.80 * W)
    O_L = lmask * lrmask
    O_L[..., :c20] = 1
    O_R = rmask * rlmask
    O_R[..., c80:] = 1
    if a_mr == 0:
        O_L, O_R = 1, 1
    rec = (rec_loss(O_R, rpan, right, vgg_right, a_p, vgg_ws) + rec_loss(O_L, lpan, left, vgg_left, a_p, vgg_ws)) / 2
    sm = 0
    if a_sm > 0:
        sm = (smoothness(left[..., c20:], ldisp[..., c20:], gamma=2) +
              smoothness(right[..., :c80], rdisp[..., :c80], gamma=2)) / 2
    mirror = 0
    if a_mr > 0:
        nmaxl = 1 / F.max_pool2d(mldisp, kernel_size=(H, W))
        nmaxr = 1 / F.max_pool2d(mrdisp, kernel_size=(H, W))
        mirror = (torch.mean(nmaxl * (1 - O_L)[..., c20:] * torch.abs(ldisp - mldisp)[..., c20:]) +
                  torch.mean(nmaxr * (1 - O_R)[..., :c80] * torch.abs(rdisp - mrdisp)[..., :c80])) / 2
    loss = rec + a_sm * sm + a_mr * mirror
    return dict(loss=loss, rec=rec, sm=sm, mirror=mirror, rpan=rpan, lpan=lpan, ldisp=ldisp, rdisp=rdisp,
                O_L=O_L, O_R=O_R)


def stage1_slow_loss(p, left, right, min_disp, max_disp, a_p=0.01, a_sm=0.2 * 2 / 512, vgg_ws=None, flip=grid_flip):
    """/root/reference/Train_Stage1_Kslow.py:236-278: Stage-2's two-view batch without masks / mirror loss."""
    B, C, H, W = left.shape
    mn2, mx2 = torch.cat((min_disp, min_disp), 0), torch.cat((max_disp, max_disp), 0)
    pan, disp = falnet_forward(p, torch.cat((left, flip(right)), 0), mn2, mx2, ret_disp=True, ret_pan=True)
    rpan, lpan = pan[:B], flip(pan[B:])
    ldisp, rdisp = disp[:B], flip(disp[B:])
    vgg_right = vgg_features(vgg_ws, right) if a_p > 0 else None
    vgg_left = vgg_features(vgg_ws, left) if a_p > 0 else None
    rec = (rec_loss(1, rpan, right, vgg_right, a_p, vgg_ws) + rec_loss(1, lpan, left, vgg_left, a_p, vgg_ws)) / 2
    c20, c80 = int(0.20 * W), int(0.80 * W)
    sm = 0
    if a_sm > 0:
        sm = (smoothness(left[..., c20:], ldisp[..., c20:], gamma=2) +
              smoothness(right[..., :c80], rdisp[..., :c80], gamma=2)) / 2
    return dict(loss=rec + a_sm * sm, rec=rec, sm=sm, rpan=rpan, lpan=lpan, ldisp=ldisp, rdisp=rdisp)


def test_disp_fpp(p, image, min_disp, max_disp, flip=None):
    """Test_KITTI flip post-processing (/root/reference/Test_KITTI.py:196-203).  The reference flips
    with grid_sample(align_corners=False defaults); an exact index flip is the intended op."""
    flip = flip or (lambda t: grid_flip(t, align_corners=False))
    d = falnet_forward(p, image, min_disp, max_disp)
    fd = flip(falnet_forward(p, flip(image), min_disp, max_disp))
    return (d + fd) / 2


def test_disp_mspp(p, image, min_disp, max_disp, flip=None):
    """Multi-scale post-processing (/root/reference/Test_KITTI.py:287-300)."""
    flip = flip or (lambda t: grid_flip(t, align_corners=False))
    B, C, H, W = image.shape
    d = falnet_forward(p, image, min_disp, max_disp)
    small = F.interpolate(flip(image), scale_factor=2 / 3, mode="bilinear", align_corners=True)
    ds = falnet_forward(p, small, min_disp, max_disp)
    ds = (1 / (2 / 3)) * F.interpolate(ds, size=(H, W), mode="nearest")
    ds = flip(ds)
    # the reference evaluates with batch 1 (Test_KITTI.py:113), so its np.percentile is PER IMAGE; for B > 1 the same
    # meaning is kept image by image
    out = []
    for b in range(B):
        db = d[b:b + 1]
        norm = db / (np.percentile(db.detach().cpu().numpy(), 95) + 1e-6)
        norm = torch.clamp(norm, max=1.0)
        out.append((1 - norm) * db + norm * ds[b:b + 1])
    return torch.cat(out, 0)


def adam_step(params, grads, m, v, step, lr, beta1=0.5, beta2=0.999, eps=1e-8):
    """torch.optim.Adam (no amsgrad, wd 0) as the entry points configure it
    (/root/reference/Train_Stage1_K.py:177-181): betas=(momentum 0.5, beta 0.999)."""
    bc1 = 1 - beta1 ** step
    bc2 = 1 - beta2 ** step
    for k in params:
        if grads.get(k) is None:
            continue
        g = grads[k]
        m[k].mul_(beta1).add_(g, alpha=1 - beta1)
        v[k].mul_(beta2).addcmul_(g, g, value=1 - beta2)
        denom = (v[k].sqrt() / math.sqrt(bc2)).add_(eps)
        params[k].addcdiv_(m[k], denom, value=-lr / bc1)


# ----------------------------------------------------------------------------------------------
# Validation metrics (numpy, like the reference)
# ----------------------------------------------------------------------------------------------
KITTI_ERROR_NAMES = ['abs_rel', 'sq_rel', 'rms', 'log_rms', 'a1', 'a2', 'a3']
WIDTH_TO_FOCAL = {1242: 721.5377, 1241: 718.856, 1224: 707.0493, 1238: 718.3351, 1226: 707.0912, 1280: 738.2355}
WIDTH_TO_BASELINE = {1242: 0.9982 * 0.54, 1241: 0.9848 * 0.54, 1224: 1.0144 * 0.54, 1238: 0.9847 * 0.54,
                     1226: 0.9765 * 0.54, 1280: 0.54}


def kitti_errors(gt, pred, min_d=1.0, max_d=80.0):
    """/root/reference/myUtils.py:196-232 (use_median=False).  gt, pred: numpy depth maps."""
    mask = gt > 0
    gt = gt[mask]
    pred = pred[mask]
    pred = np.clip(pred, min_d, max_d)
    gt = np.clip(gt, min_d, max_d)
    thresh = np.maximum(gt / pred, pred / gt)
    a1, a2, a3 = (thresh < 1.25).mean(), (thresh < 1.25 ** 2).mean(), (thresh < 1.25 ** 3).mean()
    rmse = np.sqrt(((gt - pred) ** 2).mean())
    rmse_log = np.sqrt(((np.log(gt) - np.log(pred)) ** 2).mean())
    abs_rel = np.mean(np.abs(gt - pred) / gt)
    sq_rel = np.mean(((gt - pred) ** 2) / gt)
    return [abs_rel, sq_rel, rmse, rmse_log, a1, a2, a3]


def depths_kitti2015(gt_disp, pred_disp):
    """/root/reference/myUtils.py:234-254 for one image (numpy [H,W])."""
    width = gt_disp.shape[1]
    gt_mask, pred_mask = gt_disp > 0, pred_disp > 0
    gt_depth = WIDTH_TO_FOCAL[width] * 0.54 / (gt_disp + (1.0 - gt_mask))
    pred_depth = WIDTH_TO_FOCAL[width] * 0.54 / (pred_disp + (1.0 - pred_mask))
    return gt_mask * gt_depth, pred_depth


def depths_kitti_eigen(gt_depth, pred_disp):
    """/root/reference/myUtils.py:256-277 for one image: Eigen crop, gt already a depth map."""
    height, width = gt_depth.shape
    gt = gt_depth[height - 219:height - 4, 44:1180]
    pr = pred_disp[height - 219:height - 4, 44:1180]
    gt_mask, pred_mask = gt > 0, pr > 0
    pred_depth = WIDTH_TO_FOCAL[width] * WIDTH_TO_BASELINE[width] / (pr + (1.0 - pred_mask))
    return gt_mask * gt, pred_depth


def get_rmse(output_right, label_right, mean=_RGB_MEAN):
    """/root/reference/myUtils.py:138-150."""
    m = torch.tensor(mean, dtype=output_right.dtype).view(1, 3, 1, 1)
    o = torch.clamp((output_right + m) * 255, 0, 255)
    l = (label_right + m) * 255
    return torch.mean((o - l) ** 2) ** 0.5


def real_epe(output, target, sparse=False):
    """/root/reference/loss_functions.py:124-141,170-173 for 1-channel disparity maps."""
    h, w = target.shape[2:]
    up = F.interpolate(output, size=(h, w), mode="bilinear", align_corners=True)
    epe = torch.norm(target - up, p=2, dim=1)
    if sparse:
        epe = epe[~(target[:, 0] == 0)]
    return epe.mean()
