"""tcgen05 weight-gradient kernel (csrc/conv_wgrad.cu) vs the fp32 wgrad of the same bf16 operands (torch autograd of
F.conv2d), i.e. the bound is accumulation-order noise only.  Replaces the cuDNN wgrad behind
/root/reference/models/FAL_netB.py:99-127 in loss.backward()."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

CL = torch.channels_last


def _ref_wgrad(x16, g16, cout, cin, stride):
    w = torch.zeros(cout, cin, 3, 3, device=x16.device, dtype=torch.float32, requires_grad=True)
    y = F.conv2d(x16.float()[:, :cin], w, None, stride, 1)
    (gw,) = torch.autograd.grad(y, w, g16.float()[:, :cout])
    return gw


CASES = [
    # B, H, W, Cxs, Cg, cout, cx, stride
    (2, 12, 40, 64, 64, 64, 64, 1),
    (2, 12, 40, 32, 32, 32, 32, 1),
    (2, 12, 40, 32, 64, 49, 32, 1),
    (2, 12, 40, 64, 32, 32, 64, 1),
    (1, 7, 23, 128, 64, 64, 128, 1),
    (2, 9, 20, 64, 128, 128, 64, 1),
    (1, 6, 20, 256, 256, 256, 256, 1),
    (2, 3, 10, 512, 512, 512, 512, 1),
    (2, 13, 41, 32, 64, 64, 32, 2),
    (2, 12, 40, 64, 128, 128, 64, 2),
    (1, 6, 20, 256, 512, 512, 256, 2),
    (8, 48, 160, 64, 64, 64, 64, 1),
]


@pytest.mark.parametrize("with_bias", [True, False])
@pytest.mark.parametrize("flags", [0, 1])
@pytest.mark.parametrize("B,H,W,Cxs,Cg,cout,cx,stride", CASES + [(8, 192, 640, 32, 32, 32, 32, 1)])
def test_wgrad_matches_fp32(B, H, W, Cxs, Cg, cout, cx, stride, flags, with_bias):
    """flags 0: production path (halo tile for stride-1 layers: 64-channel blocks pair two taps per MMA row, 32-channel blocks
    -- without a bias gradient -- group the taps by column); 1: nine boxes per chunk everywhere."""
    from fal_net_b200 import conv_native as CN
    torch.backends.cudnn.allow_tf32 = False
    dev = torch.device("cuda:0")
    g = torch.Generator(device=dev).manual_seed(B * 1000 + H * 10 + Cxs + stride)
    Hg, Wg = (H - 1) // stride + 1, (W - 1) // stride + 1
    x16 = torch.randn(B, Cxs, H, W, device=dev, generator=g).to(torch.bfloat16).contiguous(memory_format=CL)
    g16 = torch.randn(B, Cg, Hg, Wg, device=dev, generator=g).to(torch.bfloat16).contiguous(memory_format=CL)
    ref = _ref_wgrad(x16, g16, cout, cx, stride)
    # write into columns [8, 8 + cx) of a wider gradient tensor that already holds something (accumulation semantics)
    dW = torch.full((cout, cx + 16, 3, 3), 0.5, device=dev).contiguous(memory_format=CL)
    # ... and take the bias gradient along (the kernel's spare operand slot reads ones): accumulated into dbias[:cout]
    dbias = torch.full((cout + 3,), 2.0, device=dev) if with_bias else None
    CN.conv3x3_wgrad(g16, x16, dW, cout=cout, cx=cx, ci_off=8, stride=stride, flags=flags, dbias=dbias)
    torch.cuda.synchronize()
    if with_bias:
        bref = g16.float()[:, :cout].sum(dim=(0, 2, 3))
        assert float((dbias[:cout] - 2.0 - bref).abs().max() / bref.abs().max().clamp_min(1.0)) < 1e-3
        assert float((dbias[cout:] - 2.0).abs().max()) == 0
    got = dW[:, 8:8 + cx] - 0.5
    scale = ref.abs().max()
    assert float((got - ref).abs().max() / scale) < 2e-3, float((got - ref).abs().max() / scale)
    assert float((dW[:, :8] - 0.5).abs().max()) == 0 and float((dW[:, 8 + cx:] - 0.5).abs().max()) == 0


@pytest.mark.parametrize("stride,H,W", [(2, 12, 40), (2, 13, 41), (1, 9, 20)])
def test_const_channel_wgrad(stride, H, W):
    from fal_net_b200 import conv_native as CN
    dev = torch.device("cuda:0")
    B, C = 3, 64
    gen = torch.Generator(device=dev).manual_seed(5)
    Hg, Wg = (H - 1) // stride + 1, (W - 1) // stride + 1
    g16 = torch.randn(B, C, Hg, Wg, device=dev, generator=gen).to(torch.bfloat16).contiguous(memory_format=CL)
    val = torch.tensor([3.0, 1.5, 2.25], device=dev)
    plane = val.view(B, 1, 1, 1).expand(B, 1, H, W).contiguous()
    w = torch.zeros(C, 1, 3, 3, device=dev, requires_grad=True)
    (ref,) = torch.autograd.grad(F.conv2d(plane, w, None, stride, 1), w, g16.float())
    got = CN.const_channel_wgrad(g16, val, (H, W), stride, C)
    assert float((got - ref[:, 0]).abs().max() / ref.abs().max()) < 1e-4


UP2_CASES = [
    # B, H, W (low resolution), Cin, Cout
    (2, 6, 20, 64, 64),
    (1, 5, 17, 128, 64),          # ragged chunks: 5 rows, 17 columns
    (2, 3, 10, 512, 256),
    (3, 12, 40, 256, 128),
    (8, 48, 160, 128, 64),
    (8, 96, 320, 64, 64),
]


@pytest.mark.parametrize("B,H,W,Cin,Cout", UP2_CASES)
def test_folded_deconv_wgrad_matches_fp32(B, H, W, Cin, Cout):
    """Weight gradient of F.interpolate(scale 2, nearest) + conv3x3 (reference models/FAL_netB.py:51-60) from the LOW-resolution
    input (conv3x3_wgrad_up2_kernel: sixteen quarter-resolution correlations folded into nine taps) against fp32 autograd of
    the reference formulation on the same bf16 operands, and against the plain kernel run on the up-sampled map."""
    from fal_net_b200 import conv_native as CN
    torch.backends.cudnn.allow_tf32 = False
    dev = torch.device("cuda:0")
    gen = torch.Generator(device=dev).manual_seed(B * 100 + H + Cin)
    h16 = torch.randn(B, Cin, H, W, device=dev, generator=gen).to(torch.bfloat16).contiguous(memory_format=CL)
    g16 = torch.randn(B, Cout, 2 * H, 2 * W, device=dev, generator=gen).to(torch.bfloat16).contiguous(memory_format=CL)
    w = torch.zeros(Cout, Cin, 3, 3, device=dev, dtype=torch.float32, requires_grad=True)
    y = F.conv2d(F.interpolate(h16.float(), scale_factor=2, mode="nearest"), w, None, 1, 1)
    (ref,) = torch.autograd.grad(y, w, g16.float())
    dW = torch.full((Cout, Cin + 16, 3, 3), 0.25, device=dev).contiguous(memory_format=CL)
    CN.conv3x3_wgrad_up2(g16, h16, dW, cout=Cout, cx=Cin, ci_off=8)
    torch.cuda.synchronize()
    got = dW[:, 8:8 + Cin] - 0.25
    scale = ref.abs().max()
    assert float((got - ref).abs().max() / scale) < 2e-3, float((got - ref).abs().max() / scale)
    assert float((dW[:, :8] - 0.25).abs().max()) == 0 and float((dW[:, 8 + Cin:] - 0.25).abs().max()) == 0
    plain = torch.zeros(Cout, Cin, 3, 3, device=dev).contiguous(memory_format=CL)
    CN.conv3x3_wgrad(g16, CN.upsample_nearest(h16, (2 * H, 2 * W)), plain, cout=Cout, cx=Cin)
    assert float((got - plain).abs().max() / scale) < 2e-3


@pytest.mark.parametrize("B,H,W", [(2, 19, 45), (1, 7, 300), (8, 192, 640), (1, 5, 1242)])
def test_stem_weight_and_bias_gradient_from_the_image(B, H, W):
    """stem_wgrad_mma_kernel: dW / dbias of the 3 -> 32 first layer from the fp32 NCHW image (patch rows as bf16 hi + lo pairs in
    shared memory, the accumulator kept in TMEM across the tiles of a persistent CTA) vs fp32 autograd on the same operands;
    accumulation semantics, ragged last tile, several tiles per row."""
    from fal_net_b200 import conv_native as CN
    torch.backends.cudnn.allow_tf32 = False
    dev = torch.device("cuda:0")
    gen = torch.Generator(device=dev).manual_seed(B * 31 + W)
    x = torch.randn(B, 3, H, W, device=dev, generator=gen)
    g16 = torch.randn(B, 32, H, W, device=dev, generator=gen).to(torch.bfloat16).contiguous(memory_format=CL)
    w = torch.zeros(32, 3, 3, 3, device=dev, requires_grad=True)
    (ref,) = torch.autograd.grad(F.conv2d(x, w, None, 1, 1), w, g16.float())
    bref = g16.float().sum(dim=(0, 2, 3))
    dW = torch.full((32, 3, 3, 3), 0.5, device=dev).contiguous(memory_format=CL)
    db = torch.full((32,), -1.0, device=dev)
    CN.stem_wgrad(x, g16, dW, db)
    torch.cuda.synchronize()
    assert float((dW - 0.5 - ref).abs().max() / ref.abs().max()) < 1e-3, float((dW - 0.5 - ref).abs().max() / ref.abs().max())
    assert float((db + 1.0 - bref).abs().max() / bref.abs().max().clamp_min(1.0)) < 1e-3


def test_batched_small_map_wgrad_equals_separate_launches():
    """faln_conv3x3_wgrad_multi: the small-map layers' weight (+ bias) gradients as one grid per kernel configuration (every CTA
    finds its job by its block index) against one launch per job; mixed configurations (64 / 32-channel blocks, stride 1 / 2,
    two sources of a concatenated input, a bias gradient) in one call."""
    from fal_net_b200 import conv_native as CN
    dev = torch.device("cuda:0")
    gen = torch.Generator(device=dev).manual_seed(77)

    def rnd(*shape):
        return torch.randn(*shape, device=dev, generator=gen).to(torch.bfloat16).contiguous(memory_format=CL)
    B = 8
    specs = [  # (cin sources, cout, H, W, stride, bias)
        ((256, 256), 256, 6, 20, 1, True),      # iconv6: two sources
        ((512,), 512, 3, 10, 1, False),         # conv6_1
        ((512,), 512, 3, 10, 1, False),
        ((256,), 256, 6, 20, 1, False),         # conv5_1
        ((256,), 256, 12, 40, 1, False),        # conv4_1
        ((256,), 512, 6, 20, 2, True),          # conv6.0 (stride 2)
        ((256,), 256, 12, 40, 2, True),         # conv5.0
        ((128, 256), 256, 12, 40, 1, True),     # iconv5
        ((32,), 32, 12, 40, 1, False),          # a 32-channel block
    ]
    jobs, refs = [], []
    for srcs, cout, H, W, stride, bias in specs:
        Hg, Wg = (H - 1) // stride + 1, (W - 1) // stride + 1
        g = rnd(B, cout, Hg, Wg)
        cin = sum(srcs)
        dW_a = torch.zeros(cout, cin, 3, 3, device=dev).contiguous(memory_format=CL)
        dW_b = torch.zeros(cout, cin, 3, 3, device=dev).contiguous(memory_format=CL)
        db_a = torch.zeros(cout, device=dev) if bias else None
        db_b = torch.zeros(cout, device=dev) if bias else None
        off = 0
        for c in srcs:
            x = rnd(B, c, H, W)
            jobs.append(dict(g=g, x=x, dW=dW_a, cout=cout, cx=c, ci_off=off, stride=stride, dbias=db_a if off == 0 else None))
            CN.conv3x3_wgrad(g, x, dW_b, cout=cout, cx=c, ci_off=off, stride=stride, dbias=db_b if off == 0 else None)
            off += c
        refs.append((dW_a, dW_b, db_a, db_b))
    CN.conv3x3_wgrad_multi(jobs)
    torch.cuda.synchronize()
    for dW_a, dW_b, db_a, db_b in refs:
        scale = float(dW_b.abs().max())
        assert scale > 0 and float((dW_a - dW_b).abs().max()) / scale < 2e-3          # split-K order differs
        if db_a is not None:
            assert float((db_a - db_b).abs().max()) / float(db_b.abs().max().clamp_min(1.0)) < 1e-3


def test_batched_folded_deconv_wgrad_equals_separate_launches():
    """faln_conv3x3_wgrad_up2_multi: several folded deconv layers' weight gradients as one grid against one launch per layer."""
    from fal_net_b200 import conv_native as CN
    dev = torch.device("cuda:0")
    gen = torch.Generator(device=dev).manual_seed(78)

    def rnd(*shape):
        return torch.randn(*shape, device=dev, generator=gen).to(torch.bfloat16).contiguous(memory_format=CL)
    B = 8
    jobs, refs = [], []
    for cin, cout, H, W in ((512, 256, 3, 10), (256, 128, 6, 20), (256, 128, 12, 40), (256, 128, 24, 80)):
        h, g = rnd(B, cin, H, W), rnd(B, cout, 2 * H, 2 * W)
        a = torch.zeros(cout, cin, 3, 3, device=dev).contiguous(memory_format=CL)
        b = torch.zeros(cout, cin, 3, 3, device=dev).contiguous(memory_format=CL)
        jobs.append(dict(g=g, x=h, dW=a, cout=cout))
        CN.conv3x3_wgrad_up2(g, h, b, cout=cout)
        refs.append((a, b))
    CN.conv3x3_wgrad_up2_multi(jobs)
    torch.cuda.synchronize()
    for a, b in refs:
        scale = float(b.abs().max())
        assert scale > 0 and float((a - b).abs().max()) / scale < 2e-3
