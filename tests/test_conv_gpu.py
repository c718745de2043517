"""GPU: the tcgen05 implicit-GEMM convolution against torch conv2d on the same bf16 operands (fp32 reference math).
Tolerance: 2e-2 relative (BASELINE north_star, bf16 convolutions); in practice ~4e-3 (bf16 output rounding)."""
import pytest
import torch
import torch.nn.functional as F

from tests.helpers import rel_err

pytestmark = pytest.mark.gpu


def _ref(x, w, b, stride, act, res):
    y = F.conv2d(x.float(), w.float(), None if b is None else b.float(), stride, 1)
    if res is not None:
        y = y + res.float()
    if act == 1:
        y = F.elu(y)
    elif act == 2:
        y = F.relu(y)
    return y


@pytest.mark.parametrize("B,H,W,C1,C2,Cout,stride,act,bias,res", [
    (2, 24, 40, 64, 0, 64, 1, 1, True, False),       # BK=64, BN=64
    (1, 16, 32, 64, 0, 64, 1, 0, False, False),      # exact tiles, no epilogue extras
    (2, 24, 80, 128, 0, 256, 2, 1, True, False),     # stride 2, BN=256
    (2, 47, 156, 32, 0, 32, 1, 1, False, True),      # BK=32 (64B swizzle), residual + ELU, ragged tiles
    (1, 19, 33, 128, 0, 128, 1, 2, True, False),     # ReLU, odd sizes
    (2, 12, 40, 128, 256, 256, 1, 1, True, False),   # two sources (skip concat), K = 9*384
    (3, 3, 10, 512, 0, 512, 1, 1, False, True),      # bottleneck map smaller than one tile
    (2, 375 // 8, 1242 // 8, 256, 0, 256, 2, 1, True, False),
    (1, 6, 20, 256, 0, 512, 2, 1, True, False),
    # wide maps, one N block: the row-tile kernel (one halo load per 128-pixel row segment, resident weights)
    (2, 9, 640, 64, 0, 64, 1, 1, True, False),       # exact 5 segments per row
    (1, 11, 300, 64, 0, 64, 1, 1, False, True),      # ragged last segment + residual
    (2, 7, 333, 32, 0, 32, 1, 1, False, True),       # BK=32 (64B swizzle), 16 columns per epilogue warp
    (1, 5, 257, 64, 64, 64, 1, 1, True, False),      # two sources
    (1, 6, 200, 128, 0, 64, 1, 2, True, False),      # two K blocks of one source, ReLU
    # small maps at the training batch: the split-K cluster kernel (K loop shared by 2 / 4 / 8 CTAs, DSMEM reduce-scatter)
    (8, 6, 20, 256, 0, 256, 1, 1, True, True),       # conv5_1: BN 128 x 4 CTAs
    (8, 3, 10, 512, 0, 512, 1, 1, True, False),      # conv6_1: tiles of 2 images x 4 rows, 8 CTAs per tile
    (5, 2, 7, 256, 0, 256, 1, 0, False, False),      # tiles of 4 images x 2 rows, ragged batch
    (8, 12, 40, 256, 0, 256, 2, 1, True, False),     # conv5.0: stride 2 into a 6x20 map
    (8, 6, 20, 256, 256, 256, 1, 1, True, False),    # iconv6: two sources
    (3, 1, 5, 128, 0, 128, 1, 2, True, False),       # one-row map
    (8, 6, 20, 256, 0, 64, 1, 0, True, False),       # one N block of 64
    # wide maps, >= 8 rows (row kernel; the opt-in column-walk kernel runs the same shapes in test_column_walk_kernel_opt_in)
    (3, 40, 640, 32, 0, 32, 1, 1, True, True),       # BK = 32, CO = 32, residual
    (2, 33, 700, 64, 32, 64, 1, 1, True, False),     # two sources, three 32-channel K blocks (iconv1), ragged strip
    (2, 24, 320, 128, 0, 64, 1, 2, True, False),     # two 64-channel K blocks (iconv2)
    (1, 192, 640, 64, 0, 64, 1, 1, True, False),     # full-height strips cut into many chains
    (8, 16, 200, 64, 0, 64, 1, 0, False, False),     # chains that cross strips and images
    (1, 8, 192, 64, 0, 32, 1, 1, True, False),       # minimum height, CO = 32 with BK = 64
])
def test_conv3x3_bf16_nhwc(B, H, W, C1, C2, Cout, stride, act, bias, res):
    from fal_net_b200 import conv_native as CN
    dev = torch.device("cuda:0")
    g = torch.Generator().manual_seed(H * 1000 + W + C1 + Cout)
    CL = torch.channels_last
    x = torch.randn(B, C1, H, W, generator=g).bfloat16().to(dev).contiguous(memory_format=CL)
    x2 = torch.randn(B, C2, H, W, generator=g).bfloat16().to(dev).contiguous(memory_format=CL) if C2 else None
    Cin = C1 + C2
    w = (torch.randn(Cout, Cin, 3, 3, generator=g) * (2.0 / (9 * Cin)) ** 0.5).bfloat16().to(dev)
    b = torch.randn(Cout, generator=g).to(dev) if bias else None
    Ho, Wo = (H - 1) // stride + 1, (W - 1) // stride + 1
    r = torch.randn(B, Cout, Ho, Wo, generator=g).bfloat16().to(dev).contiguous(memory_format=CL) if res else None
    y = CN.conv3x3_fwd(x, CN.pack_weight(w), b, stride, act, r, x2)
    torch.cuda.synchronize()
    xin = x if x2 is None else torch.cat((x, x2), 1)
    ref = _ref(xin, w, b, stride, act, r)
    assert y.shape == ref.shape
    e = rel_err(y.float(), ref)
    assert e < 2e-2, e
    assert e < 8e-3, e


def test_conv3x3_planar_fp32_logits():
    """Last layer: concat(64 + 32) -> 49 planes, fp32 planar output with a padded row pitch."""
    from fal_net_b200 import conv_native as CN, layout
    dev = torch.device("cuda:0")
    g = torch.Generator().manual_seed(3)
    CL = torch.channels_last
    B, H, W, N = 2, 21, 70, 49
    u = torch.randn(B, 64, H, W, generator=g).bfloat16().to(dev).contiguous(memory_format=CL)
    s = torch.randn(B, 32, H, W, generator=g).bfloat16().to(dev).contiguous(memory_format=CL)
    w = (torch.randn(N, 96, 3, 3, generator=g) * 0.05).bfloat16().to(dev)
    b = torch.randn(N, generator=g).to(dev)
    out = layout.alloc_planar(B, N, H, W, dev)
    assert out.stride(2) == 72
    CN.conv3x3_fwd(u, CN.pack_weight(w, 64), b, 1, 0, None, s, cout=N, planar_out=out)
    torch.cuda.synchronize()
    ref = _ref(torch.cat((u, s), 1), w, b, 1, 0, None)
    assert rel_err(out, ref) < 1e-3          # fp32 straight from the accumulator: only operand rounding is shared
    # the same layer on a wide map goes through the row-tile kernel (BK = 32, three channel blocks, planar stores)
    B, H, W = 1, 6, 270
    u = torch.randn(B, 64, H, W, generator=g).bfloat16().to(dev).contiguous(memory_format=CL)
    s = torch.randn(B, 32, H, W, generator=g).bfloat16().to(dev).contiguous(memory_format=CL)
    out = layout.alloc_planar(B, N, H, W, dev)
    CN.conv3x3_fwd(u, CN.pack_weight(w, 64), b, 1, 0, None, s, cout=N, planar_out=out)
    torch.cuda.synchronize()
    assert rel_err(out, _ref(torch.cat((u, s), 1), w, b, 1, 0, None)) < 1e-3


def test_stem_upsample_pool_and_const_channel():
    from fal_net_b200 import conv_native as CN
    dev = torch.device("cuda:0")
    g = torch.Generator().manual_seed(9)
    CL = torch.channels_last
    x = torch.randn(2, 3, 19, 45, generator=g).to(dev)
    for Cout, act in ((32, 1), (64, 2)):
        w = (torch.randn(Cout, 3, 3, 3, generator=g) * 0.3).to(dev)
        b = torch.randn(Cout, generator=g).to(dev)
        y = CN.stem_conv(x, w, b, act)
        ref = _ref(x, w, b, 1, act, None)
        assert rel_err(y.float(), ref) < 8e-3
        # the tensor-core form (FALN_STEM_TC=1): image patch and weights are bf16 operands like every other layer
        CN.STEM_TC = True
        try:
            yt = CN.stem_conv(x, w, b, act)
            yf = CN.stem_conv(x, w, b, act, flip_x=True)
        finally:
            CN.STEM_TC = False
        ref16 = _ref(x.bfloat16().float(), w.bfloat16().float(), b, 1, act, None)
        assert rel_err(yt.float(), ref16) < 8e-3 and rel_err(yt.float(), ref) < 2e-2
        assert rel_err(yf.float(), _ref(torch.flip(x, dims=[3]).bfloat16().float(), w.bfloat16().float(), b, 1, act, None)) < 8e-3
    # nearest upsample, non-integer ratios like the 375x1242 pyramid (24 -> 47, 78 -> 156)
    a = torch.randn(2, 64, 24, 78, generator=g).bfloat16().to(dev).contiguous(memory_format=CL)
    up = CN.upsample_nearest(a, (47, 156))
    assert torch.equal(up, F.interpolate(a, size=(47, 156), mode="nearest"))
    up2 = CN.upsample_nearest(a, (48, 156))
    assert torch.equal(up2, F.interpolate(a, size=(48, 156), mode="nearest"))
    assert torch.equal(CN.maxpool2(a), F.max_pool2d(a, 2, 2))
    odd = torch.randn(1, 64, 23, 77, generator=g).bfloat16().to(dev).contiguous(memory_format=CL)
    assert torch.equal(CN.maxpool2(odd), F.max_pool2d(odd, 2, 2))
    # constant extra input channel folded into a border-class bias (conv1.0: 32 + 1 -> 64, stride 2), even and odd sizes
    for H, W in ((24, 40), (25, 41)):
        h = torch.randn(2, 32, H, W, generator=g).bfloat16().to(dev).contiguous(memory_format=CL)
        w = (torch.randn(64, 33, 3, 3, generator=g) * 0.1).bfloat16().to(dev)
        b = torch.randn(64, generator=g).to(dev)
        cval = torch.tensor([3.0, 1.25], device=dev)
        y = CN.conv3x3_fwd(h, CN.pack_weight(w[:, :32]), b, 2, 1, None, None, ctab=CN.const_channel_table(w[:, 32].float()),
                           cscale=cval)
        plane = cval.view(2, 1, 1, 1).expand(2, 1, H, W).to(torch.bfloat16)
        ref = _ref(torch.cat((h, plane), 1), w, b, 2, 1, None)
        assert rel_err(y.float(), ref) < 8e-3


def _elu_grad_from_y(y):
    return torch.where(y > 0, torch.ones_like(y), y + 1)


@pytest.mark.parametrize("B,H,W,Cin,Cout,stride,dact,accum,res", [
    (2, 24, 40, 64, 64, 1, 1, False, False),
    (2, 47, 156, 32, 32, 1, 1, False, True),       # residual pass-through + ELU' (res block)
    (1, 19, 33, 128, 128, 1, 2, False, False),     # ReLU' (VGG)
    (2, 24, 80, 128, 256, 2, 1, True, False),      # stride 2, even sizes, accumulate into an existing gradient
    (2, 25, 81, 64, 128, 2, 1, False, False),      # stride 2, odd sizes
    (2, 47, 155, 256, 256, 2, 0, False, False),    # stride 2, mixed parity, no activation
    (3, 3, 10, 512, 512, 1, 1, False, False),
    (2, 21, 70, 96, 49, 1, 0, False, False),       # logits conv: Cout 49 padded to 64 on K
    (2, 9, 640, 64, 64, 1, 1, False, False),       # row-tile kernel (wide map, one N block)
    (1, 7, 333, 32, 32, 1, 1, False, True),
    (1, 6, 260, 64, 49, 1, 0, False, False),
    # split-K cluster kernel
    (8, 6, 20, 256, 256, 1, 1, False, True),
    (8, 12, 40, 256, 256, 2, 1, True, False),      # stride 2 (four tap classes of 1 / 2 / 2 / 4 taps), accumulate
    (5, 3, 10, 512, 512, 1, 1, False, False),
    (8, 11, 39, 256, 256, 2, 0, False, False),     # stride 2, odd sizes
    (8, 6, 20, 512, 256, 1, 0, False, False),      # concat split: two row ranges of the re-packed weights
    # more wide-map shapes
    (2, 40, 640, 32, 32, 1, 1, False, True),
    (2, 33, 320, 64, 64, 1, 1, True, False),       # accumulate into an existing gradient
    (1, 50, 300, 64, 49, 1, 0, False, False),      # logits conv: Cout 49 padded to 64 on K
    (3, 17, 200, 32, 64, 1, 2, False, False),      # 64 -> 32 channels, ReLU'
])
def test_conv3x3_dgrad(B, H, W, Cin, Cout, stride, dact, accum, res):
    from fal_net_b200 import conv_native as CN
    dev = torch.device("cuda:0")
    gen = torch.Generator().manual_seed(H * 1000 + W + Cin + Cout)
    CL = torch.channels_last
    Ho, Wo = (H - 1) // stride + 1, (W - 1) // stride + 1
    Cg = (Cout + 31) // 32 * 32
    w = (torch.randn(Cout, Cin, 3, 3, generator=gen) * (2.0 / (9 * Cout)) ** 0.5).bfloat16().to(dev)
    g = torch.zeros(B, Cg, Ho, Wo)
    g[:, :Cout] = torch.randn(B, Cout, Ho, Wo, generator=gen)
    g = g.bfloat16().to(dev).contiguous(memory_format=CL)
    ysave = torch.randn(B, Cin, H, W, generator=gen).bfloat16().to(dev).contiguous(memory_format=CL) if dact else None
    old = torch.randn(B, Cin, H, W, generator=gen).bfloat16().to(dev).contiguous(memory_format=CL) if accum else None
    r = torch.randn(B, Cin, H, W, generator=gen).bfloat16().to(dev).contiguous(memory_format=CL) if res else None
    ref = torch.nn.grad.conv2d_input((B, Cin, H, W), w.float(), g[:, :Cout].float(), stride, 1)
    if old is not None:
        ref = ref + old.float()
    if r is not None:
        ref = ref + r.float()
    if dact == 1:
        ref = ref * _elu_grad_from_y(ysave.float())
    elif dact == 2:
        ref = ref * (ysave.float() > 0)
    out = old.clone(memory_format=CL) if old is not None else None
    wd = CN.pack_weight_dgrad(w)
    if Cin % 64 == 0 and not accum and not res and out is None:
        # two row ranges written into two tensors = the concat split
        h = Cin // 2
        a = CN.conv3x3_dgrad(g, wd, (H, W), stride, rows=(0, h), dact=dact, ysave=None if ysave is None else ysave[:, :h])
        b = CN.conv3x3_dgrad(g, wd, (H, W), stride, rows=(h, h), dact=dact, ysave=None if ysave is None else ysave[:, h:])
        got = torch.cat((a, b), 1)
    else:
        got = CN.conv3x3_dgrad(g, wd, (H, W), stride, out=out, accum=accum, dact=dact, ysave=ysave, residual=r)
    torch.cuda.synchronize()
    e = rel_err(got.float(), ref)
    assert e < 8e-3, e


@pytest.mark.parametrize("B,H,W", [(2, 19, 45), (1, 7, 300), (2, 12, 640), (1, 5, 1242)])
def test_fused_tensor_core_stem(B, H, W):
    """stem_mma_kernel (default stem: 27-tap patch rows built in shared memory as bf16 hi + lo pairs, four tcgen05 MMAs per 128
    pixels) vs fp32 torch on the fp32 image with bf16-rounded weights -- the image itself must NOT be rounded to bf16 -- and vs
    the fp32-FMA kernel it replaces; ragged last tile, several tiles per row, horizontal flip, both widths / activations."""
    from fal_net_b200 import conv_native as CN
    dev = torch.device("cuda:0")
    g = torch.Generator().manual_seed(B * 7 + W)
    x = torch.randn(B, 3, H, W, generator=g).to(dev)
    assert CN.STEM_MMA
    for Cout, act in ((32, 1), (64, 2), (32, 0)):
        w = (torch.randn(Cout, 3, 3, 3, generator=g) * 0.3).to(dev)
        b = torch.randn(Cout, generator=g).to(dev)
        y = CN.stem_conv(x, w, b, act)
        ref_w16 = _ref(x, w.bfloat16().float(), b, 1, act, None)
        e = rel_err(y.float(), ref_w16)
        assert e < 5e-3, (Cout, act, e)                                   # bf16 output rounding only (2^-9 of the max)
        if act == 0:
            # without the output's own bf16 rounding in the way: a bf16-rounded image would show ~4e-3 here
            yy = y.float()
            near = (yy - ref_w16).abs() <= (ref_w16.abs() * 2 ** -8 + 1e-3)
            assert bool(near.all())
        yf = CN.stem_conv(x, w, b, act, flip_x=True)
        assert rel_err(yf.float(), _ref(torch.flip(x, dims=[3]), w.bfloat16().float(), b, 1, act, None)) < 5e-3
        CN.STEM_MMA = False
        try:
            y_fma = CN.stem_conv(x, w, b, act)
        finally:
            CN.STEM_MMA = True
        assert rel_err(y.float(), y_fma.float()) < 8e-3


def test_upsample_pool_backward_and_channel_sum():
    from fal_net_b200 import conv_native as CN
    dev = torch.device("cuda:0")
    gen = torch.Generator().manual_seed(11)
    CL = torch.channels_last
    for (Hl, Wl), (Hh, Wh) in (((24, 78), (47, 156)), ((3, 10), (6, 20)), ((188, 621), (375, 1242)), ((12, 40), (12, 40))):
        C = 64 if Hh < 300 else 8
        x = torch.randn(2, C, Hl, Wl, generator=gen).bfloat16().to(dev).contiguous(memory_format=CL)
        gh = torch.randn(2, C, Hh, Wh, generator=gen).bfloat16().to(dev).contiguous(memory_format=CL)
        xr = x.float().requires_grad_(True)
        F.interpolate(xr, size=(Hh, Wh), mode="nearest").backward(gh.float())
        ref = xr.grad * _elu_grad_from_y(x.float())
        got = CN.upsample_nearest_bwd(gh, (Hl, Wl), ysave=x, dact=1)
        assert rel_err(got.float(), ref) < 8e-3
    # max-pool backward with ReLU' (distinct values: no ties)
    x = torch.randn(2, 64, 23, 78, generator=gen).bfloat16().to(dev).contiguous(memory_format=CL)
    gy = torch.randn(2, 64, 11, 39, generator=gen).bfloat16().to(dev).contiguous(memory_format=CL)
    xr = x.float().requires_grad_(True)
    F.max_pool2d(xr, 2, 2).backward(gy.float())
    ref = xr.grad * (x.float() > 0)
    assert rel_err(CN.maxpool2_bwd(x, gy, dact=2).float(), ref) < 1e-6
    # ... with ties (ReLU outputs are full of equal zeros, bf16 values collide): the first maximum in scan order gets the
    # gradient, as in ATen's max_pool2d_with_indices; even and odd sizes, several channel counts
    for (Hh, Wh, C) in ((24, 80, 64), (13, 31, 128), (6, 20, 256)):
        x = (torch.randn(2, C, Hh, Wh, generator=gen).clamp_min(0) * 4).round().div(4).bfloat16().to(dev).contiguous(memory_format=CL)
        gy = torch.randn(2, C, Hh // 2, Wh // 2, generator=gen).bfloat16().to(dev).contiguous(memory_format=CL)
        xr = x.float().requires_grad_(True)
        F.max_pool2d(xr, 2, 2).backward(gy.float())
        ref = xr.grad * (x.float() > 0)
        got = CN.maxpool2_bwd(x, gy, dact=2).float()
        assert torch.equal(got, ref), (Hh, Wh, C, float((got - ref).abs().max()))
    # channel sums (bias gradient), accumulating
    for C, Cs in ((64, 64), (49, 64), (256, 256), (32, 32), (512, 512)):
        g = torch.randn(2, Cs, 30, 50, generator=gen).bfloat16().to(dev).contiguous(memory_format=CL)
        out = torch.ones(C, device=dev)
        CN.channel_sum(g, out, C)
        ref = 1 + g.float().sum((0, 2, 3))[:C]
        assert rel_err(out, ref) < 1e-4


def test_vgg_slices_forward_and_input_gradient():
    """The hand-scheduled VGG node (fal_net_b200.conv: tcgen05 forward, dgrad-only backward with fused ReLU' / pool routing /
    slice-gradient adds) vs torch autograd through fp32 conv2d / relu / max_pool2d
    (/root/reference/loss_functions.py:21-29,36-44).  The reference rounds every activation to bf16 (straight-through), so
    both sides take the ReLU / arg-max decisions on the same values; what is left is the bf16 rounding of the gradient
    tensors themselves (2^-9 per layer): 3e-2 in rel-L2."""
    import torch.nn.functional as F
    from fal_net_b200 import loss_functions as LF
    dev = torch.device("cuda:0")
    vgg = LF.Vgg19_pc().to(dev)
    g = torch.Generator().manual_seed(3)
    x = (torch.rand(2, 3, 48, 80, generator=g) - 0.43).to(dev).requires_grad_(True)
    outs = vgg(x)
    cots = [torch.randn(o.shape, generator=g).to(dev) for o in outs]
    loss = sum((o.float() * c).sum() for o, c in zip(outs, cots))
    (gx,) = torch.autograd.grad(loss, x)

    xr = x.detach().clone().requires_grad_(True)
    h, i, refs = xr, 0, []
    from fal_net_b200.conv import VGG_CFG
    for v in VGG_CFG:
        if v == "M":
            h = F.max_pool2d(h, 2, 2)
            refs.append(h)
        else:
            w16 = vgg.weights[i].to(torch.bfloat16).float() if i else vgg.weights[i].float()
            h = F.relu(F.conv2d(h, w16, vgg.biases[i].float(), 1, 1))
            h = h + (h.to(torch.bfloat16).float() - h).detach()          # bf16 activation storage, straight-through
            i += 1
    lref = sum((o * c).sum() for o, c in zip(refs, cots))
    (gr,) = torch.autograd.grad(lref, xr)
    for o, r in zip(outs, refs):
        assert float((o.float() - r).abs().max() / r.abs().max()) < 2e-2
    assert gx.shape == gr.shape and gx.dtype == torch.float32
    # yardstick: the same slices through cuDNN's bf16 convolutions (torch autocast).  A ReLU / arg-max decision taken on an
    # activation one bf16 ulp away re-routes a gradient discretely, so the error of ANY bf16 implementation against the fp32
    # chain is a few 1e-2 here and moves with the summation order; the rule is the one of tests/test_reference_gpu.py:
    # ours <= max(2e-2, 1.5 x cuDNN-bf16).
    xa = x.detach().clone().requires_grad_(True)
    h, i, outs_a = xa, 0, []
    with torch.autocast("cuda", dtype=torch.bfloat16):
        for v in VGG_CFG:
            if v == "M":
                h = F.max_pool2d(h, 2, 2)
                outs_a.append(h)
            else:
                h = F.relu(F.conv2d(h, vgg.weights[i], vgg.biases[i], 1, 1))
                i += 1
    (ga,) = torch.autograd.grad(sum((o.float() * c).sum() for o, c in zip(outs_a, cots)), xa)
    yard = float((ga - gr).norm() / gr.norm())
    e = float((gx - gr).norm() / gr.norm())
    assert e < max(2e-2, 1.5 * yard), (e, yard)


@pytest.mark.parametrize("B,Cin,Cout,H,W", [(2, 64, 64, 24, 40), (1, 256, 128, 6, 10), (2, 128, 64, 13, 21), (1, 512, 256, 3, 10),
                                            (2, 64, 64, 96, 320),
                                            (8, 512, 256, 3, 10), (8, 256, 128, 6, 20), (3, 256, 128, 5, 9)])  # split-K cluster kernel
def test_upsample_folded_conv_forward_and_data_gradient(B, Cin, Cout, H, W):
    """deconv block (nearest 2x up-sampling + conv3x3 + ELU, reference :51-60) as ONE kernel with folded weights: forward
    against fp32 F.interpolate + conv2d on the same bf16 operands, data gradient (w.r.t. the LOW-resolution input, times the
    producer's ELU') against fp32 autograd; and both against the explicit two-kernel path of the product."""
    import torch.nn.functional as F
    from fal_net_b200 import conv_native as CN
    dev = torch.device("cuda:0")
    g = torch.Generator(device=dev).manual_seed(B * 100 + Cin + H)
    x = torch.randn(B, Cin, H, W, generator=g, device=dev).to(torch.bfloat16).contiguous(memory_format=torch.channels_last)
    w = (torch.randn(Cout, Cin, 3, 3, generator=g, device=dev) / (3 * Cin ** 0.5))
    wf, wd = CN.pack_up2_weights(w)
    y = CN.conv3x3_up2_fwd(x, wf, None, 1)
    xr = x.float().requires_grad_(True)
    w16 = w.to(torch.bfloat16).float()
    yr = F.elu(F.conv2d(F.interpolate(xr, scale_factor=2, mode="nearest"), w16, None, 1, 1))
    assert y.shape == yr.shape
    # folded weights are bf16(sum of fp32 taps), the reference multiplies bf16-rounded taps: within the bf16 bound
    assert float((y.float() - yr).abs().max() / yr.abs().max()) < 2e-2
    y2 = CN.conv3x3_fwd(CN.upsample_nearest(x, (2 * H, 2 * W)), CN.pack_weight(w), None, 1, 1)
    assert float((y.float() - y2.float()).abs().max() / y2.float().abs().max()) < 2e-2
    # data gradient with the producer's ELU' (ysave = an ELU output, i.e. > -1)
    gy = torch.randn(B, Cout, 2 * H, 2 * W, generator=g, device=dev).to(torch.bfloat16).contiguous(memory_format=torch.channels_last)
    pre = torch.randn(B, Cin, H, W, generator=g, device=dev)
    ys = F.elu(pre).to(torch.bfloat16).contiguous(memory_format=torch.channels_last)
    gx = CN.conv3x3_up2_dgrad(gy, wd, dact=1, ysave=ys)
    lin = F.conv2d(F.interpolate(xr, scale_factor=2, mode="nearest"), w16, None, 1, 1)
    (gr,) = torch.autograd.grad((lin * gy.float()).sum(), xr)
    ysf = ys.float()
    gr = gr * torch.where(ysf > 0, torch.ones_like(ysf), ysf + 1)
    assert gx.shape == gr.shape
    assert float((gx.float() - gr).abs().max() / gr.abs().max()) < 2e-2


def test_vgg_full_returns_the_fourth_slice():
    """Vgg19_pc(x, full=True) (reference loss_functions.py:36-44): four pooled activations, the fourth after conv4_1..4_4."""
    from fal_net_b200 import loss_functions as LF
    dev = torch.device("cuda:0")
    vgg = LF.Vgg19_pc().to(dev)
    x = (torch.rand(2, 3, 64, 96, generator=torch.Generator().manual_seed(3)) - 0.43).to(dev)
    with torch.no_grad():
        outs = vgg(x, full=True)
    assert len(outs) == 4 and tuple(outs[3].shape) == (2, 512, 4, 6)
    h = outs[2].float()
    for w, b in zip(vgg.weights4, vgg.biases4):
        h = F.relu(F.conv2d(h.bfloat16().float(), w.bfloat16().float(), b, 1, 1))
    ref = F.max_pool2d(h, 2, 2)
    assert rel_err(outs[3].float(), ref) < 2e-2


@pytest.mark.parametrize("H,W,stride", [(24, 40, 2), (25, 41, 2), (24, 40, 1), (13, 21, 1)])
def test_const_channel_weight_gradient_against_autograd(H, W, stride):
    """the weight gradient of the spatially constant max_disp/100 input plane (border-class sums, one kernel to
    combine them) against autograd of a conv over an EXPLICIT constant plane -- stride 2 with even and odd sizes, stride 1."""
    from fal_net_b200 import conv_native as CN
    dev = torch.device("cuda:0")
    g = torch.Generator(device=dev).manual_seed(H * 7 + W)
    B, Cout = 3, 64
    Hg, Wg = (H - 1) // stride + 1, (W - 1) // stride + 1
    gy = torch.randn(B, Cout, Hg, Wg, generator=g, device=dev).to(torch.bfloat16).contiguous(memory_format=torch.channels_last)
    val = torch.tensor([3.0, 1.25, 0.5], device=dev)
    w = torch.zeros(Cout, 1, 3, 3, device=dev, requires_grad=True)
    plane = val.view(B, 1, 1, 1).expand(B, 1, H, W).contiguous()
    y = F.conv2d(plane, w, None, stride, 1)
    (ref,) = torch.autograd.grad((y * gy.float()).sum(), w)
    for krsc in (False, True):                                 # plain NCHW gradient tensor and the arena's KRSC layout
        dW = torch.zeros(Cout, 3, 3, 33, device=dev).permute(0, 3, 1, 2) if krsc else torch.zeros(Cout, 33, 3, 3, device=dev)
        CN.const_channel_wgrad_into(gy, val, (H, W), stride, Cout, dW, 32)
        assert rel_err(dW[:, 32], ref[:, 0]) < 1e-5
        assert float(dW[:, :32].abs().max()) == 0.0


_COL_SCRIPT = r"""
import torch, torch.nn.functional as F
from fal_net_b200 import conv_native as CN
from tests.helpers import rel_err
dev = torch.device("cuda:0"); CL = torch.channels_last
g = torch.Generator().manual_seed(5)
for (B, H, W, C1, C2, Cout, act, res) in [(3, 40, 640, 32, 0, 32, 1, True), (2, 33, 700, 64, 32, 64, 1, False), (1, 192, 640, 64, 0, 64, 1, False),
                                           (8, 16, 200, 64, 0, 64, 0, False), (1, 8, 192, 64, 0, 32, 1, False), (2, 9, 640, 64, 0, 64, 1, False)]:
    x = torch.randn(B, C1, H, W, generator=g).bfloat16().to(dev).contiguous(memory_format=CL)
    x2 = torch.randn(B, C2, H, W, generator=g).bfloat16().to(dev).contiguous(memory_format=CL) if C2 else None
    Cin = C1 + C2
    w = (torch.randn(Cout, Cin, 3, 3, generator=g) * (2.0 / (9 * Cin)) ** 0.5).bfloat16().to(dev)
    b = torch.randn(Cout, generator=g).to(dev)
    r = torch.randn(B, Cout, H, W, generator=g).bfloat16().to(dev).contiguous(memory_format=CL) if res else None
    y = CN.conv3x3_fwd(x, CN.pack_weight(w), b, 1, act, r, x2)
    xin = x if x2 is None else torch.cat((x, x2), 1)
    ref = F.conv2d(xin.float(), w.float(), b, 1, 1)
    if r is not None: ref = ref + r.float()
    if act == 1: ref = F.elu(ref)
    e = rel_err(y.float(), ref); assert e < 8e-3, ("fwd", B, H, W, e)
    if C2 == 0:
        gy = torch.randn(B, Cout, H, W, generator=g).bfloat16().to(dev).contiguous(memory_format=CL)
        ys = torch.randn(B, Cin, H, W, generator=g).bfloat16().to(dev).contiguous(memory_format=CL)
        got = CN.conv3x3_dgrad(gy, CN.pack_weight_dgrad(w), (H, W), 1, dact=1, ysave=ys)
        gr = torch.nn.grad.conv2d_input((B, Cin, H, W), w.float(), gy.float(), 1, 1)
        gr = gr * torch.where(ys.float() > 0, torch.ones_like(gr), ys.float() + 1)
        e = rel_err(got.float(), gr); assert e < 8e-3, ("dgrad", B, H, W, e)
torch.cuda.synchronize(); print("COL_OK")
"""


def test_column_walk_kernel_opt_in():
    """The column-walk kernel (FALN_CONV_COL=1; off by default, csrc/conv_tc.cu) is read once per process, so its parity check runs
    in a child process: six wide-map shapes forward (bias / ELU / residual / two sources) and data gradient (ELU') against fp32
    torch on the same bf16 operands, and the debug log must show that the kernel was the one launched."""
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ, FALN_CONV_COL="1", FALN_DEBUG="1", PYTHONPATH=root)
    r = subprocess.run([sys.executable, "-c", _COL_SCRIPT], env=env, cwd=root, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and "COL_OK" in r.stdout, r.stderr[-2000:]
    assert "conv3x3_col_kernel" in r.stderr
