"""FAL_netA / FAL_netC and the Kslow step.

CPU: the product's constructors reproduce the reference's initial weights (tests/golden/variants.npz, generated from the
reference's own models/FAL_netA.py / FAL_netC.py); the oracle's variant tables reproduce the reference's forward (golden).
GPU: forward of the product against the golden reference outputs and the oracle; Stage-1 gradients of every used tensor
against oracle autograd; the Kslow step against the oracle."""
import numpy as np
import pytest
import torch

from oracle import falnet_oracle as O
from tests.helpers import disp_range, images, rel_err, rel_l2

VARIANTS = ("FAL_netA", "FAL_netC")


@pytest.fixture(scope="module")
def gold(golden_dir):
    return np.load(f"{golden_dir}/variants.npz")


@pytest.mark.parametrize("name", VARIANTS)
def test_product_constructor_matches_reference_stream(gold, name):
    from fal_net_b200 import models
    torch.manual_seed(0)
    m = models.__dict__[name](None)
    sd = m.state_dict()
    N = int(gold[f"{name}_levels"])
    assert m.no_levels == N and list(sd.keys()) == list(O.param_shapes(N, name).keys())
    assert np.array_equal(np.array([float(v.double().sum()) for v in sd.values()]), gold[f"{name}_init_sums"])
    assert np.array_equal(np.array([float(v.double().abs().sum()) for v in sd.values()]), gold[f"{name}_init_abs"])
    m2 = models.__dict__[name]({"state_dict": sd}, no_levels=N)              # checkpoint contract
    assert all(torch.equal(a, b) for a, b in zip(m2.state_dict().values(), sd.values()))
    n_used = len(m.used_parameters())
    assert n_used == len(sd) - (3 if name != "FAL_netA" else 0)


@pytest.mark.parametrize("name", VARIANTS)
def test_oracle_variant_forward_matches_reference_golden(gold, name):
    from fal_net_b200 import models
    B, H, W = (int(v) for v in gold["meta"])
    torch.manual_seed(0)
    p = {k: v.clone() for k, v in models.__dict__[name](None).state_dict().items()}
    left = images(B, H, W, 1234)
    mn, mx = disp_range(B)
    with torch.no_grad():
        o = O.falnet_forward(p, left, mn, mx, True, True, True)
    for t, nm in zip(o, ("pan", "disp", "maskL", "maskR")):
        assert np.array_equal(t.numpy(), gold[f"{name}_{nm}"]), (name, nm)


@pytest.mark.gpu
@pytest.mark.parametrize("name", VARIANTS)
def test_variant_forward_on_device(gold, name):
    from fal_net_b200 import models
    dev = torch.device("cuda:0")
    B, H, W = (int(v) for v in gold["meta"])
    torch.manual_seed(0)
    m = models.__dict__[name](None).to(dev)
    N = m.no_levels
    p = {k: v.detach().cpu().clone() for k, v in m.state_dict().items()}
    left = images(B, H, W, 1234)
    mn, mx = disp_range(B)
    with torch.no_grad():
        pan, disp, mL, mR = m(left.to(dev), mn.to(dev), mx.to(dev), ret_disp=True, ret_subocc=True, ret_pan=True)
        donly = m(left.to(dev), mn.to(dev), mx.to(dev))
        lg = m.logits(left.to(dev), mx.to(dev))[..., :W].cpu().contiguous()
    # (1) conv outputs (logits) within the bf16 bound, max-norm
    flow = torch.ones(B, 1, H, W) * (mx.view(B, 1, 1, 1) / 100)
    rlg = torch.nn.functional.conv2d(O.backbone_forward(p, left, flow), p["conv0.weight"], p["conv0.bias"])
    assert rel_err(lg, rlg) < 2e-2 and rel_l2(lg, rlg) < 2e-2
    # (2) MED section given THOSE logits: fp32 bound (incl. FAL_netA's align_corners=False maskR)
    ref = O.med_forward_ops(lg, left, mn, mx, True, True, True, maskr_align=O.VARIANTS[name]["maskr_align"])
    for a, b, nm in zip((pan, disp, mL, mR), ref, ("pan", "disp", "maskL", "maskR")):
        assert rel_err(a, b) < 1e-4, (name, nm, rel_err(a, b))
    # (3) end to end against the reference's golden outputs
    for a, nm in zip((pan, disp, mL, mR), ("pan", "disp", "maskL", "maskR")):
        assert rel_l2(a, torch.from_numpy(gold[f"{name}_{nm}"])) < 2e-2, (name, nm)
    assert rel_l2(donly, torch.from_numpy(gold[f"{name}_disp"])) < 2e-2


@pytest.mark.gpu
@pytest.mark.parametrize("name", VARIANTS)
def test_variant_stage1_gradients_all_tensors(name):
    from fal_net_b200 import models, steps
    dev = torch.device("cuda:0")
    torch.manual_seed(0)
    m = models.__dict__[name](None).to(dev)
    p = {k: v.detach().cpu().clone() for k, v in m.state_dict().items()}
    B, H, W = 2, 64, 192
    left, right = images(B, H, W, 1234), images(B, H, W, 1235)
    mn, mx = disp_range(B)
    loss = steps.stage1_loss(m, left.to(dev), right.to(dev), mn.to(dev), mx.to(dev), a_p=0.0)[0]
    loss.backward()
    pp = {k: v.clone().requires_grad_(True) for k, v in p.items()}
    lo = O.stage1_loss(pp, left, right, mn, mx, a_p=0.0)[0]
    lo.backward()
    assert rel_err(loss, lo) < 1e-2
    for k, q in m.used_parameters():
        assert q.grad.shape == pp[k].grad.shape
        e = rel_l2(q.grad, pp[k].grad)
        assert e < 8e-2, (name, k, e)


@pytest.mark.gpu
def test_variant_trains_through_flat_adam():
    """FAL_netA through the flat-arena optimiser (3x1 / 1x3 weights live un-packed in the arena): the loss moves and every
    parameter stays finite."""
    from fal_net_b200 import models, steps
    from fal_net_b200.trainer import FlatAdamDDP
    dev = torch.device("cuda:0")
    torch.manual_seed(0)
    m = models.FAL_netA(None).to(dev)
    opt = FlatAdamDDP(m, lr=1e-3)
    B, H, W = 2, 64, 192
    left, right = images(B, H, W, 1).to(dev), images(B, H, W, 2).to(dev)
    mn, mx = (t.to(dev) for t in disp_range(B))
    losses = []
    for _ in range(4):
        opt.zero_grad()
        loss = steps.stage1_loss(m, left, right, mn, mx, a_p=0.0)[0]
        loss.backward()
        opt.step()
        losses.append(float(loss))
    assert losses[-1] < losses[0] and torch.isfinite(opt.p).all()


@pytest.mark.gpu
def test_kslow_step_matches_oracle():
    from fal_net_b200 import loss_functions as LF, models, steps
    import torchvision
    dev = torch.device("cuda:0")
    torch.manual_seed(0)
    m = models.FAL_netB(None).to(dev)
    p = {k: v.detach().cpu().clone() for k, v in m.state_dict().items()}
    B, H, W = 2, 48, 160
    left, right = images(B, H, W, 1234), images(B, H, W, 1235)
    mn, mx = disp_range(B)
    torch.manual_seed(2)
    sd = torchvision.models.vgg19().state_dict()
    ws = [(sd[f"features.{i}.weight"], sd[f"features.{i}.bias"]) for i in (0, 2, 5, 7, 10, 12, 14, 16)]
    vgg = LF.Vgg19_pc().to(dev)
    res = steps.stage1_slow_loss(m, left.to(dev), right.to(dev), mn.to(dev), mx.to(dev), a_p=0.01, vgg=vgg)
    res["loss"].backward()
    pp = {k: v.clone().requires_grad_(True) for k, v in p.items()}
    ref = O.stage1_slow_loss(pp, left, right, mn, mx, a_p=0.01, vgg_ws=ws, flip=lambda t: torch.flip(t, dims=[3]))
    ref["loss"].backward()
    for k, tol in (("loss", 1e-2), ("rec", 1e-2), ("sm", 3e-2)):
        assert rel_err(res[k], ref[k]) < tol, (k, float(res[k]), float(ref[k]))
    for k, q in m.used_parameters():
        assert rel_l2(q.grad, pp[k].grad) < 8e-2, k
