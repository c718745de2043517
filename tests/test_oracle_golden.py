"""CPU: the oracle reproduces the committed reference outputs (tests/golden/*.npz, generated from the
imported reference by tests/golden/make_golden.py)."""
import os

import numpy as np
import torch

from oracle import falnet_oracle as O
from tests.helpers import disp_range, images, med_case_inputs, rel_err, sha


def test_param_table_matches_reference_constructor(golden_dir):
    g = np.load(os.path.join(golden_dir, "init_checksums.npz"))
    for N in (49, 33):
        shapes = O.param_shapes(N)
        assert len(shapes) == len(g[f"init{N}_sums"]) == 50
        assert sum(int(np.prod(s)) for s in shapes.values()) == int(g[f"init{N}_nparam"])
    assert int(g["init49_nparam"]) == 16974354


def test_med_ops_oracle_matches_golden(golden_dir):
    g = np.load(os.path.join(golden_dir, "med_cases.npz"))
    for tag in "abcdef":
        B, N, H, W = (int(v) for v in g[f"{tag}_meta"])
        logits, img, gp, gd = med_case_inputs(tag, B, N, H, W)
        assert sha(logits) + sha(img) + sha(gp) + sha(gd) == str(g[f"{tag}_in_sha"]), "seeded inputs drifted"
        mn, mx = torch.from_numpy(g[f"{tag}_min"]), torch.from_numpy(g[f"{tag}_max"])
        logits.requires_grad_(True)
        pan, disp, mL, mR = O.med_forward_ops(logits, img, mn, mx, True, True, True)
        (gl,) = torch.autograd.grad((pan * gp).sum() + (disp * gd).sum(), logits)
        for nm, t in (("pan", pan), ("disp", disp), ("maskL", mL), ("maskR", mR)):
            assert rel_err(t, torch.from_numpy(g[f"{tag}_{nm}"])) < 2e-6, (tag, nm)
        assert rel_err(gl, torch.from_numpy(g[f"{tag}_glogits"])) < 1e-5, tag


def test_med_closed_form_matches_golden(golden_dir):
    """The gather formulation the CUDA kernels implement == the reference's grid_sample path."""
    g = np.load(os.path.join(golden_dir, "med_cases.npz"))
    for tag in "abcdef":
        B, N, H, W = (int(v) for v in g[f"{tag}_meta"])
        logits, img, gp, gd = med_case_inputs(tag, B, N, H, W)
        mn, mx = torch.from_numpy(g[f"{tag}_min"]), torch.from_numpy(g[f"{tag}_max"])
        d, xo = O.level_tables(mn, mx, N, W)
        cl = O.med_forward_closed(logits, img, d, xo)
        for nm in ("pan", "disp", "maskL", "maskR"):
            assert rel_err(cl[nm], torch.from_numpy(g[f"{tag}_{nm}"])) < 5e-6, (tag, nm)
        gl = O.med_backward_closed(logits, img, d, xo, gp, gd)
        assert rel_err(gl, torch.from_numpy(g[f"{tag}_glogits"])) < 5e-6, tag


def _vgg_ws():
    import torchvision
    torch.manual_seed(2)
    sd = torchvision.models.vgg19().state_dict()
    return [(sd[f"features.{i}.weight"], sd[f"features.{i}.bias"]) for i in (0, 2, 5, 7, 10, 12, 14, 16)]


def test_vgg_standin_is_reproducible(golden_dir):
    g = np.load(os.path.join(golden_dir, "vgg_seed2_checksums.npz"))
    sums = np.array([float(w.double().sum()) for w, _ in _vgg_ws()])
    assert np.allclose(sums, g["sums"], rtol=0, atol=1e-9)


def test_losses_match_golden(golden_dir):
    g = np.load(os.path.join(golden_dir, "net_small.npz"))
    B, N, H, W = (int(v) for v in g["meta"])
    left, right = images(B, H, W, 1234), images(B, H, W, 1235)
    pan, disp = torch.from_numpy(g["fwd_pan"]), torch.from_numpy(g["fwd_disp"])
    ws = _vgg_ws()
    assert rel_err(O.rec_loss(1, pan, right, None, 0), torch.from_numpy(g["loss_rec_plain"])) < 1e-6
    msk = torch.from_numpy(g["mask_seed5"])
    r2 = O.rec_loss(msk, pan, right, O.vgg_features(ws, right), 0.01, ws)
    assert rel_err(r2, torch.from_numpy(g["loss_rec_masked_vgg"])) < 1e-5
    c0 = int(0.2 * W)
    s1 = O.smoothness(left[..., c0:], disp[..., c0:], gamma=2)
    assert rel_err(s1, torch.from_numpy(g["loss_smooth"])) < 1e-6


def test_adam_matches_torch():
    torch.manual_seed(3)
    p0 = {"a": torch.randn(37), "b": torch.randn(5, 4)}
    ref = {k: v.clone().requires_grad_(True) for k, v in p0.items()}
    opt = torch.optim.Adam(list(ref.values()), lr=1e-4, betas=(0.5, 0.999))
    mine = {k: v.clone() for k, v in p0.items()}
    m = {k: torch.zeros_like(v) for k, v in p0.items()}
    v = {k: torch.zeros_like(w) for k, w in p0.items()}
    for step in range(1, 4):
        grads = {k: torch.randn_like(w) for k, w in p0.items()}
        for k in ref:
            ref[k].grad = grads[k].clone()
        opt.step()
        O.adam_step(mine, grads, m, v, step, 1e-4)
        for k in ref:
            assert torch.allclose(ref[k].detach(), mine[k], rtol=1e-6, atol=1e-7)


def test_product_level_tables_bit_identical_to_reference_loop():
    """fal_net_b200.med.level_tables_torch (the vectorised torch form the device kernel is checked against on the GPU,
    tests/test_med_gpu.py) == the reference's per-level expressions, bit for bit (CPU)."""
    from fal_net_b200 import med
    for N in (9, 33, 49, 65):
        for W in (160, 640, 1242, 2048):
            for mx, mn in ((300.0, 2.0), (60.0, 0.75), (192.3, 1.7)):
                mxx = torch.tensor([mx, mx * 0.7]).view(2, 1, 1)
                mnn = torch.tensor([mn, mn * 1.3]).view(2, 1, 1)
                d0, x0 = O.level_tables(mnn, mxx, N, W)
                d1, x1 = med.level_tables_torch(mnn, mxx, N, W)
                assert torch.equal(d0, d1) and torch.equal(x0, x1)


def test_product_constructor_draws_the_reference_random_stream(golden_dir):
    """torch.manual_seed(0); FAL_netB(no_levels=N) of the product gives the reference's initial weights
    (/root/reference/models/FAL_netB.py:130-138,190-192): per-tensor sums and abs-sums recorded from the reference."""
    import numpy as np
    import torch
    from fal_net_b200 import models
    g = np.load(f"{golden_dir}/init_checksums.npz")
    for N in (49, 33):
        torch.manual_seed(0)
        sd = models.FAL_netB(no_levels=N).state_dict()
        assert int(g[f"init{N}_nparam"]) == sum(v.numel() for v in sd.values())
        sums = np.array([float(v.double().sum()) for v in sd.values()])
        abss = np.array([float(v.double().abs().sum()) for v in sd.values()])
        assert np.array_equal(sums, g[f"init{N}_sums"]) and np.array_equal(abss, g[f"init{N}_abs"])
