"""GPU: the whole FAL_netB path (bf16 convolutions + fused MED + fused losses) against the CPU oracle with the
same weights and inputs.  Tolerances are BASELINE.json north_star's: bf16 convolutions <= 2e-2 relative on
outputs, <= 1e-2 on losses."""
import pytest
import torch

from oracle import falnet_oracle as O
from tests.helpers import disp_range, images, rel_err, rel_l2

pytestmark = pytest.mark.gpu
OUT_TOL, LOSS_TOL = 2e-2, 1e-2
# End-to-end outputs of a RANDOM-INIT network (a softmax over 49 nearly-flat logits of magnitude ~15) are the worst case
# for logit noise.  Here, against the CPU oracle at small shapes: the conv outputs (logits) are held to 2e-2 in max-norm,
# the MED section given those logits to 1e-4, and the end-to-end outputs to 2e-2 in relative L2.  The end-to-end
# MAX-norm is asserted at BASELINE shapes in tests/test_reference_gpu.py::test_whole_model_at_baseline_shapes against
# the reference on the same GPU, with the reference's own cuDNN-bf16 run as the yardstick (no hand-widened constants).


def _dev():
    assert torch.cuda.is_available()
    return torch.device("cuda:0")


def _model(seed=0, N=49):
    from fal_net_b200 import models
    torch.manual_seed(seed)
    m = models.__dict__["FAL_netB"](None, no_levels=N).to(_dev())
    p = {k: v.detach().cpu().clone() for k, v in m.state_dict().items()}
    return m, p


def _vgg_ws():
    import torchvision
    torch.manual_seed(2)
    sd = torchvision.models.vgg19().state_dict()
    return [(sd[f"features.{i}.weight"], sd[f"features.{i}.bias"]) for i in (0, 2, 5, 7, 10, 12, 14, 16)]


def test_state_dict_contract():
    m, p = _model()
    assert list(p.keys()) == list(O.param_shapes(49).keys())
    assert len(m.weight_parameters()) == 36 and len(m.bias_parameters()) == 14
    from fal_net_b200 import models
    m2 = models.FAL_netB({"state_dict": p}, no_levels=49)          # checkpoint dict contract (reference :28-32)
    assert all(torch.equal(a, b) for a, b in zip(m2.state_dict().values(), p.values()))


@pytest.mark.parametrize("H,W", [(48, 160), (50, 166)])
def test_forward_all_outputs(H, W):
    dev = _dev()
    m, p = _model()
    B = 2
    left = images(B, H, W, 1234)
    mn, mx = disp_range(B)
    with torch.no_grad():
        pan, disp, mL, mR = m(left.to(dev), mn.to(dev), mx.to(dev), ret_disp=True, ret_subocc=True, ret_pan=True)
        donly = m(left.to(dev), mn.to(dev), mx.to(dev))
    rp, rd, rmL, rmR = O.falnet_forward(p, left, mn, mx, True, True, True)
    assert isinstance(donly, torch.Tensor) and rel_l2(donly, rd) < OUT_TOL
    assert rel_l2(disp, rd) < OUT_TOL
    # (1) the convolution outputs themselves (the logits) meet the bf16 bound in max-norm
    with torch.no_grad():
        lg = m.logits(left.to(dev), mx.to(dev))[..., :W].cpu().contiguous()
    flow = torch.ones(B, 1, H, W) * (mx.view(B, 1, 1, 1) / 100)
    rlg = torch.nn.functional.conv2d(O.backbone_forward(p, left, flow), p["conv0.weight"], p["conv0.bias"])
    assert rel_err(lg, rlg) < OUT_TOL and rel_l2(lg, rlg) < OUT_TOL
    # (2) given THOSE logits, the fused MED kernels reproduce the reference's synthesis to the fp32 bound
    d, xo = O.level_tables(mn, mx, 49, W)
    ref = O.med_forward_closed(lg, left, d, xo)
    assert rel_err(pan, ref["pan"]) < 1e-4 and rel_err(disp, ref["disp"]) < 1e-4
    assert rel_err(mL, ref["maskL"]) < 1e-4 and rel_err(mR, ref["maskR"]) < 1e-4
    # (3) end to end, relative L2 (max-norm: see the note at the top of this file)
    assert rel_l2(pan, rp) < OUT_TOL and rel_l2(mL, rmL) < OUT_TOL and rel_l2(mR, rmR) < OUT_TOL


def test_stage1_step_loss_and_gradients():
    from fal_net_b200 import steps
    dev = _dev()
    m, p = _model()
    B, H, W = 2, 48, 160
    left, right = images(B, H, W, 1234), images(B, H, W, 1235)
    mn, mx = disp_range(B)
    loss, rec, sm, _, _ = steps.stage1_loss(m, left.to(dev), right.to(dev), mn.to(dev), mx.to(dev), a_p=0.0)
    loss.backward()
    pp = {k: v.clone().requires_grad_(True) for k, v in p.items()}
    lo, rrec, rsm, _, _ = O.stage1_loss(pp, left, right, mn, mx, a_p=0.0)
    lo.backward()
    assert rel_err(loss, lo) < LOSS_TOL and rel_err(rec, rrec) < LOSS_TOL and rel_err(sm, rsm) < 3e-2
    named = dict(m.named_parameters())
    # every used tensor against the oracle's autograd (fp32 CPU): relative L2.  bf16 activation / gradient storage puts
    # the decoder at ~1e-2 and the far end of the 34-layer chain at 4e-2 ... 6e-2 at this tiny shape (2x48x160; cuDNN bf16
    # shows the same, tests/test_reference_gpu.py); a wrong scale, a missing term or a mis-indexed tap in any
    # hand-scheduled gradient shows up as >= 1e-1.  The per-tensor comparison against the
    # reference's own cuDNN-bf16 backward is tests/test_reference_gpu.py::test_stage1_parameter_gradients_all_tensors.
    n_checked = 0
    for k, q in named.items():
        if "amask_conv" in k:
            continue
        e = rel_l2(q.grad, pp[k].grad)
        assert e < 8e-2, (k, e)
        n_checked += 1
    assert n_checked == 47
    assert named["backbone.amask_conv.0.weight"].grad is None


def test_stage1_with_perceptual_loss():
    from fal_net_b200 import loss_functions as LF, steps
    dev = _dev()
    m, p = _model()
    B, H, W = 2, 48, 160
    left, right = images(B, H, W, 1234), images(B, H, W, 1235)
    mn, mx = disp_range(B)
    ws = _vgg_ws()
    vgg = LF.Vgg19_pc().to(dev)
    for a, (w, _) in zip(vgg.weights, ws):
        assert torch.equal(a.detach().cpu(), w)
    loss = steps.stage1_loss(m, left.to(dev), right.to(dev), mn.to(dev), mx.to(dev), a_p=0.01, vgg=vgg)[0]
    lo = O.stage1_loss(p, left, right, mn, mx, a_p=0.01, vgg_ws=ws)[0]
    assert rel_err(loss, lo) < LOSS_TOL


def test_stage2_loss():
    from fal_net_b200 import loss_functions as LF, steps
    dev = _dev()
    m, p = _model(0)
    mf, pf = _model(1)
    B, H, W = 2, 48, 160
    left, right = images(B, H, W, 1234), images(B, H, W, 1235)
    mn, mx = disp_range(B)
    ws = _vgg_ws()
    vgg = LF.Vgg19_pc().to(dev)
    res = steps.stage2_loss(m, mf, left.to(dev), right.to(dev), mn.to(dev), mx.to(dev), a_p=0.01, vgg=vgg)
    res["loss"].backward()
    # the oracle with exact index flips injected on both sides
    ref = O.stage2_loss(p, pf, left, right, mn, mx, a_p=0.01, vgg_ws=ws, flip=lambda t: torch.flip(t, dims=[3]))
    for k, tol in (("loss", LOSS_TOL), ("rec", LOSS_TOL), ("sm", 3e-2), ("mirror", 3e-2)):
        assert rel_err(res[k], ref[k]) < tol, (k, float(res[k]), float(ref[k]))
    assert rel_l2(res["O_L"], ref["O_L"]) < OUT_TOL and rel_l2(res["O_R"], ref["O_R"]) < OUT_TOL
    # gradients of every used tensor (mirror loss, flip-folded rec / smoothness, VGG dgrad into the backbone)
    pp = {k: v.clone().requires_grad_(True) for k, v in p.items()}
    O.stage2_loss(pp, pf, left, right, mn, mx, a_p=0.01, vgg_ws=ws, flip=lambda t: torch.flip(t, dims=[3]))["loss"].backward()
    for k, q in m.used_parameters():
        e = rel_l2(q.grad, pp[k].grad)
        assert e < 8e-2, (k, e)


def test_inference_post_processing():
    from fal_net_b200 import steps
    dev = _dev()
    m, p = _model()
    B, H, W = 1, 54, 180
    img = images(B, H, W, 77)
    mn, mx = disp_range(B)
    flip = lambda t: torch.flip(t, dims=[3])
    d_fpp = steps.test_disp(m, img.to(dev), mn.to(dev), mx.to(dev), f_post_process=True)
    r_fpp = O.test_disp_fpp(p, img, mn, mx, flip=flip)
    assert rel_l2(d_fpp, r_fpp) < OUT_TOL
    d_ms = steps.test_disp(m, img.to(dev), mn.to(dev), mx.to(dev), ms_post_process=True)
    r_ms = O.test_disp_mspp(p, img, mn, mx, flip=flip)
    assert rel_l2(d_ms, r_ms) < OUT_TOL


def test_graphed_step_matches_eager():
    """The CUDA-graph replay of a whole Stage-1 step (zero_grad, forward, losses, backward, Adam with the device-side step
    counter) walks the same parameter trajectory as launching the step kernel by kernel."""
    from fal_net_b200 import steps
    from fal_net_b200.trainer import FlatAdamDDP, GraphedStep
    dev = _dev()
    B, H, W = 2, 48, 160
    batches = [(images(B, H, W, 10 + i).to(dev), images(B, H, W, 20 + i).to(dev)) for i in range(4)]
    mn, mx = (t.to(dev) for t in disp_range(B))

    def run(graph):
        m, _ = _model()
        opt = FlatAdamDDP(m, lr=2e-4)
        fn = lambda l, r: steps.stage1_loss(m, l, r, mn, mx, a_p=0.0)[0]
        losses = []
        if graph:
            gs = GraphedStep(opt, fn, *batches[0], warmup=0)
            for l, r in batches:
                losses.append(float(gs.run(l, r)))
        else:
            for l, r in batches:
                opt.zero_grad()
                loss = fn(l, r)
                loss.backward()
                opt.step()
                losses.append(float(loss))
        torch.cuda.synchronize()
        return losses, opt.p.clone()

    le, pe = run(False)
    lg, pg = run(True)
    assert le[0] != le[-1]                                   # the parameters did move
    for a, b in zip(le, lg):
        assert abs(a - b) <= 5e-3 * abs(a), (le, lg)         # wgrad split-K order differs run to run (fp32 red.add)
    assert rel_l2(pg, pe) < 1e-3


def test_graphed_step_lagged_loss_readback():
    """GraphedStep.run_async / loss_value (the end-to-end loop of bench.py): the host queues step i + 1 before it reads the
    loss of step i; every ticket must return exactly what a synchronous ``run()`` of the same trajectory returns, with the
    prefetch staging in between, and a ticket that has left the ring must be refused."""
    from fal_net_b200 import steps
    from fal_net_b200.trainer import FlatAdamDDP, GraphedStep
    dev = _dev()
    B, H, W = 2, 48, 160
    host = [(images(B, H, W, 30 + i).pin_memory(), images(B, H, W, 40 + i).pin_memory()) for i in range(6)]
    mn, mx = (t.to(dev) for t in disp_range(B))

    def run(lagged):
        m, _ = _model()
        opt = FlatAdamDDP(m, lr=2e-4)
        fn = lambda l, r: steps.stage1_loss(m, l, r, mn, mx, a_p=0.0)[0]
        gs = GraphedStep(opt, fn, host[0][0].to(dev), host[0][1].to(dev), warmup=0)
        out = []
        if not lagged:
            for l, r in host:
                out.append(float(gs.run(l.to(dev), r.to(dev))))
            return out, gs
        gs.prefetch(*host[0])
        pending = None
        for i in range(len(host)):
            t = gs.run_async()
            if i + 1 < len(host):
                gs.prefetch(*host[i + 1])
            if pending is not None:
                out.append(gs.loss_value(pending))
            pending = t
        out.append(gs.loss_value(pending))
        return out, gs

    sync, _ = run(False)
    lag, gs = run(True)
    assert len(lag) == len(sync) == len(host)
    for a, b in zip(sync, lag):
        assert abs(a - b) <= 5e-3 * abs(a), (sync, lag)      # same tolerance as graph vs eager (fp32 red.add order)
    assert sync[0] != sync[-1]
    with pytest.raises(RuntimeError):
        gs.loss_value(1)                                     # six steps later the ring (depth 4) has re-used that slot


def test_fused_disparity_epilogue_matches_unfused():
    """Inference: the logits layer with the softmax-expectation fused into its epilogue (the N planes never reach HBM)
    against the same layer writing fp32 logits followed by the disparity kernel (reference :215-229)."""
    from fal_net_b200 import backbone, med
    dev = _dev()
    m, _ = _model()
    B, H, W = 2, 37, 333
    img = images(B, H, W, 5).to(dev)
    mn, mx = (t.to(dev) for t in disp_range(B))
    with torch.no_grad():
        fused = m(img, mn, mx, ret_disp=True, ret_subocc=False, ret_pan=False)
        d_lvl, _ = med.level_tables(mn, mx, m.no_levels, W)
        logits = backbone.forward(m, img, mx, None)
        unfused = med.med_disp_only(logits, d_lvl)
    assert fused.shape == unfused.shape == (B, 1, H, W)
    assert float((fused - unfused).abs().max() / unfused.abs().max()) < 1e-4
