"""GPU parity of the loss / optimiser / layout kernels (through the C ABI) against the CPU oracle."""
import pytest
import torch

from oracle import falnet_oracle as O
from tests.helpers import images, rel_err

pytestmark = pytest.mark.gpu
LOSS_TOL = 1e-4          # BASELINE north_star allows 1e-2 on losses; fp32 kernels do far better


def _dev():
    assert torch.cuda.is_available()
    return torch.device("cuda:0")


def test_rec_l1_forward_backward():
    from fal_net_b200 import losses
    dev = _dev()
    B, H, W = 3, 24, 100
    g = torch.Generator().manual_seed(1)
    synth = images(B, H, W, 1) + 0.1 * torch.randn(B, 3, H, W, generator=g)
    label = images(B, H, W, 2)
    mask = torch.rand(B, 1, H, W, generator=g)
    for m in (None, mask):
        for flip in (False, True):
            s = synth.clone().requires_grad_(True)
            s_un = torch.flip(s, dims=[3]) if flip else s
            mm = 1 if m is None else m
            ref = torch.mean(mm * torch.abs(s_un - label))
            blend_ref = mm * s_un + (1 - mm) * label
            gb = torch.randn(B, 3, H, W, generator=g)
            (ref * 0.7 + (blend_ref * gb).sum()).backward()
            val, blend = losses.rec_l1(s.detach().to(dev), label.to(dev), None if m is None else m.to(dev),
                                       want_blend=True, flip_x=flip)
            gs = losses.rec_l1_bwd(s.detach().to(dev), label.to(dev), None if m is None else m.to(dev), gb.to(dev), 0.7,
                                   flip_x=flip)
            assert rel_err(val, ref) < LOSS_TOL
            assert rel_err(blend, blend_ref) < 1e-6
            assert rel_err(gs, s.grad) < 1e-5


@pytest.mark.parametrize("lo_frac,hi_frac,flip", [(0.2, 1.0, False), (0.0, 0.8, True), (0.0, 1.0, False)])
def test_smoothness(lo_frac, hi_frac, flip):
    from fal_net_b200 import losses
    dev = _dev()
    B, H, W = 2, 20, 160
    g = torch.Generator().manual_seed(2)
    img = images(B, H, W, 4)
    disp = (50 * torch.rand(B, 1, H, W, generator=g)).requires_grad_(True)
    lo, hi = int(lo_frac * W), int(hi_frac * W)
    dd = torch.flip(disp, dims=[3]) if flip else disp
    ref = O.smoothness(img[..., lo:hi], dd[..., lo:hi], gamma=2)
    ref.backward()
    val = losses.smoothness(img.to(dev), disp.detach().to(dev), 2.0, lo, hi, flip_x=flip)
    gd = torch.zeros(B, 1, H, W, device=dev)
    losses.smoothness_bwd(img.to(dev), disp.detach().to(dev), 2.0, 1.0, gd, lo, hi, flip_x=flip)
    assert rel_err(val, ref) < LOSS_TOL
    assert rel_err(gd, disp.grad) < 1e-4


def test_mirror_and_occlusion_masks():
    from fal_net_b200 import losses
    dev = _dev()
    B, H, W = 2, 16, 80
    g = torch.Generator().manual_seed(3)
    disp = (40 * torch.rand(B, 1, H, W, generator=g)).requires_grad_(True)
    mdisp = 40 * torch.rand(B, 1, H, W, generator=g)
    a, b = torch.rand(B, 1, H, W, generator=g), torch.rand(B, 1, H, W, generator=g)
    c20, c80 = int(0.2 * W), int(0.8 * W)
    # O_L = lmask * unflip(lrmask); first 20 % := 1   (Train_Stage2_K.py:296-297)
    O_L = a * torch.flip(b, dims=[3])
    O_L[..., :c20] = 1
    o_l = losses.occ_mask(a.to(dev), b.to(dev), False, True, 0, c20)
    assert rel_err(o_l, O_L) < 1e-7
    O_R = torch.flip(a, dims=[3]) * b
    O_R[..., c80:] = 1
    o_r = losses.occ_mask(a.to(dev), b.to(dev), True, False, c80, W)
    assert rel_err(o_r, O_R) < 1e-7
    import torch.nn.functional as F
    nmax = 1 / F.max_pool2d(mdisp, kernel_size=(H, W))
    inv = losses.inv_rowmax(mdisp.to(dev))
    assert rel_err(inv.view(-1), nmax.view(-1)) < 1e-7
    for flip, lo, hi, occ in ((False, c20, W, O_L), (True, 0, c80, O_R)):
        dd = torch.flip(disp, dims=[3]) if flip else disp
        ref = torch.mean(nmax * (1 - occ)[..., lo:hi] * torch.abs(dd - mdisp)[..., lo:hi])
        disp.grad = None
        ref.backward()
        val = losses.mirror(disp.detach().to(dev), mdisp.to(dev), occ.to(dev), inv, lo, hi, flip_x=flip)
        gd = torch.zeros(B, 1, H, W, device=dev)
        losses.mirror_bwd(disp.detach().to(dev), mdisp.to(dev), occ.to(dev), inv, 1.0, gd, lo, hi, flip_x=flip)
        assert rel_err(val, ref) < LOSS_TOL
        assert rel_err(gd, disp.grad) < 1e-5


def test_mse_bf16():
    from fal_net_b200 import losses
    dev = _dev()
    g = torch.Generator().manual_seed(4)
    a = torch.randn(2, 12, 40, 64, generator=g).bfloat16()
    b = torch.randn(2, 12, 40, 64, generator=g).bfloat16()
    ref = torch.mean((a.float() - b.float()) ** 2)
    val = losses.mse_bf16(a.to(dev), b.to(dev))
    assert rel_err(val, ref) < 1e-5
    ga = losses.mse_bf16_bwd(a.to(dev), b.to(dev), 0.5)
    gref = 0.5 * 2 * (a.float() - b.float()) / a.numel()
    assert rel_err(ga.float(), gref) < 1e-2        # bf16 output rounding


def test_adam_matches_torch_and_oracle():
    from fal_net_b200 import optim
    dev = _dev()
    torch.manual_seed(0)
    n = 4096 + 8
    p0 = torch.randn(n)
    ref = p0.clone().requires_grad_(True)
    opt = torch.optim.Adam([ref], lr=1e-4, betas=(0.5, 0.999))
    p = p0.clone().to(dev)
    m, v = torch.zeros(n, device=dev), torch.zeros(n, device=dev)
    w16 = torch.empty(n, device=dev, dtype=torch.bfloat16)
    for step in range(1, 6):
        gr = torch.randn(n)
        ref.grad = gr.clone()
        opt.step()
        optim.adam_step_(p, gr.to(dev), m, v, w16, lr=1e-4, beta1=0.5, beta2=0.999, eps=1e-8, step=step)
    assert rel_err(p, ref.detach()) < 1e-6
    assert torch.equal(w16.cpu(), p.cpu().bfloat16())


def test_layout_kernels():
    from fal_net_b200 import layout
    dev = _dev()
    g = torch.Generator().manual_seed(6)
    x = torch.randn(2, 3, 9, 70, generator=g)
    for flip in (False, True):
        y = layout.nchw_to_nhwc_bf16(x.to(dev), 8, flip_x=flip)
        xr = torch.flip(x, dims=[3]) if flip else x
        ref = torch.zeros(2, 9, 70, 8)
        ref[..., :3] = xr.permute(0, 2, 3, 1)
        assert torch.equal(y.cpu().float(), ref.bfloat16().float())
    z = torch.randn(2, 49, 5, 70, generator=g)
    zn = layout.planar_to_nhwc_bf16(z.to(dev), 64)
    ref = torch.zeros(2, 5, 70, 64)
    ref[..., :49] = z.permute(0, 2, 3, 1)
    assert torch.equal(zn.cpu().float(), ref.bfloat16().float())
    back = layout.nhwc_bf16_to_planar(zn, 49, pitch=72)
    assert back.shape == (2, 49, 5, 70) and back.stride(2) == 72
    assert torch.equal(back.cpu(), z.bfloat16().float())
    # 16-byte aligned planar rows (the layout of the logits / their gradient): wide transpose kernel, incl. a ragged last
    # quad (W % 4 != 0), more than one 128-pixel block per row, and Cp = 32
    for C, Cp, W in ((49, 64, 70), (49, 64, 301), (17, 32, 130), (64, 64, 128)):
        zz = torch.randn(2, C, 3, W, generator=g)
        zp = layout.alloc_planar(2, C, 3, W, dev)
        zp.copy_(zz.to(dev))
        out = layout.planar_to_nhwc_bf16(zp, Cp)
        ref = torch.zeros(2, 3, W, Cp)
        ref[..., :C] = zz.permute(0, 2, 3, 1)
        assert torch.equal(out.cpu().float(), ref.bfloat16().float()), (C, Cp, W)


def test_flat_adam_two_param_groups_like_the_reference():
    """The reference builds Adam with two param groups -- bias_parameters with bias_decay, weight_parameters with weight_decay
    (Train_Stage1_K.py:177-181).  FlatAdamDDP with distinct decays against torch.optim.Adam on the same parameters / grads."""
    from fal_net_b200 import models
    from fal_net_b200.trainer import FlatAdamDDP
    dev = torch.device("cuda:0")
    torch.manual_seed(0)
    m = models.FAL_netB(no_levels=9).to(dev)
    ref = {n: p.detach().clone().requires_grad_(True) for n, p in m.used_parameters()}
    opt = FlatAdamDDP(m, lr=1e-3, betas=(0.5, 0.999), weight_decay=1e-2, bias_decay=0.0)
    topt = torch.optim.Adam([{"params": [p for n, p in ref.items() if "bias" in n], "weight_decay": 0.0},
                             {"params": [p for n, p in ref.items() if "weight" in n], "weight_decay": 1e-2}],
                            lr=1e-3, betas=(0.5, 0.999))
    g = torch.Generator(device=dev).manual_seed(1)
    for _ in range(3):
        opt.zero_grad()
        opt.g.copy_(torch.randn(opt.g.shape, generator=g, device=dev) * 1e-2)
        for n, p in m.used_parameters():
            ref[n].grad = p.grad.detach().clone().contiguous()
        opt.step()
        topt.step()
    for n, p in m.used_parameters():
        assert float((p.detach() - ref[n].detach()).abs().max()) < 2e-6, n


def test_flat_dgrad_repack_equals_the_two_dimensional_launch():
    """faln_pack_dgrad_flat (flat tile list, a warp per 32x32 tile, 4-byte accesses; odd row length of the 33-channel
    conv1.0 included) writes bit for bit what faln_pack_dgrad_batched writes, for every 3x3 weight of FAL_netB."""
    from fal_net_b200 import _lib, models
    from fal_net_b200.trainer import FlatAdamDDP
    dev = torch.device("cuda:0")
    torch.manual_seed(3)
    m = models.__dict__["FAL_netB"](None, no_levels=49).to(dev)
    opt = FlatAdamDDP(m, lr=1e-4)
    assert opt._total_tiles == sum(9 * (int(j[2]) // 32) * (int(j[4]) // 32) for j in opt._jobs.tolist())
    assert any(int(j[3]) % 2 == 1 for j in opt._jobs.tolist())          # the odd-pitch path is exercised
    a = torch.zeros_like(opt.wd16)
    b = torch.zeros_like(opt.wd16)
    L = _lib.lib()
    _lib.check(L.faln_pack_dgrad_batched(_lib.ptr(opt.w16), _lib.ptr(a), _lib.ptr(opt._jobs), opt._jobs.shape[0],
                                         opt._max_tiles, _lib.cur_stream()), "2d")
    _lib.check(L.faln_pack_dgrad_flat(_lib.ptr(opt.w16), _lib.ptr(b), _lib.ptr(opt._jobs), opt._jobs.shape[0],
                                      opt._total_tiles, _lib.cur_stream()), "flat")
    torch.cuda.synchronize()
    assert float(a.float().abs().sum()) > 0
    assert torch.equal(a.view(torch.int16), b.view(torch.int16))
