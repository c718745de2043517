"""Import the UNMODIFIED reference (baseline/_ref, see tools/install_reference.py) for reference-on-CUDA parity tests
and bench.py's reference legs.  Test / measurement infrastructure only: nothing under fal_net_b200/ imports this.

The reference's loss_functions.py builds ``vgg19(pretrained=True)`` at import (loss_functions.py:10,48); there is no
network, so a seeded random VGG19 state-dict (seed 2 -- the same stand-in fal_net_b200.loss_functions.Vgg19_pc builds)
is written where torch.hub looks before downloading.  On a CUDA box nothing else is shimmed; on a CPU
box (``cpu=True``) ``.cuda()`` becomes the identity, as in tests/golden/make_golden.py (process-wide: bench.py runs its CPU
reference leg in a subprocess for that reason).
"""
import os
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_DIR = os.path.join(ROOT, "baseline", "_ref")     # git-ignored, NOT gpurun-ignored: travels to the GPU box
_CACHE = {}


def available():
    return os.path.exists(os.path.join(REF_DIR, "models", "FAL_netB.py"))


def vgg_state_dict(seed=2):
    import torch
    import torchvision
    rng = torch.random.get_rng_state()
    torch.manual_seed(seed)
    sd = torchvision.models.vgg19().state_dict()
    torch.random.set_rng_state(rng)
    return sd


def load(cpu=False, want_losses=True):
    """Returns (models_module, loss_functions_module_or_None) of the reference."""
    key = (cpu, want_losses)
    if key in _CACHE:
        return _CACHE[key]
    if not available():
        raise RuntimeError(f"{REF_DIR} is missing: run `python tools/install_reference.py` in the build container")
    import torch
    import torch.nn as nn
    if cpu:
        torch.Tensor.cuda = lambda self, *a, **k: self
        nn.Module.cuda = lambda self, *a, **k: self
    if REF_DIR not in sys.path:
        sys.path.insert(0, REF_DIR)
    import models as ref_models
    assert os.path.abspath(ref_models.__file__).startswith(REF_DIR), ref_models.__file__
    ref_losses = None
    if want_losses:
        if "loss_functions" in sys.modules and os.path.abspath(sys.modules["loss_functions"].__file__).startswith(REF_DIR):
            ref_losses = sys.modules["loss_functions"]
        else:
            tmp = tempfile.mkdtemp(prefix="falnet_torchhome_")
            old = os.environ.get("TORCH_HOME")
            os.environ["TORCH_HOME"] = tmp
            ck = os.path.join(tmp, "hub", "checkpoints")
            os.makedirs(ck)
            torch.save(vgg_state_dict(), os.path.join(ck, "vgg19-dcbb9e9d.pth"))
            try:
                import warnings
                with warnings.catch_warnings():
                    warnings.simplefilter("ignore")
                    import loss_functions as ref_losses
            finally:
                if old is None:
                    os.environ.pop("TORCH_HOME", None)
                else:
                    os.environ["TORCH_HOME"] = old
                import shutil
                shutil.rmtree(tmp, ignore_errors=True)
            assert os.path.abspath(ref_losses.__file__).startswith(REF_DIR)
    _CACHE[key] = (ref_models, ref_losses)
    return _CACHE[key]


# ---------------------------------------------------------------------------------------------------------------
# The step bodies of the reference's train loops, calling the reference's own modules in the reference's order.
# The scripts themselves need tensorboardX / KITTI on disk so their loop bodies are restated here.
# ---------------------------------------------------------------------------------------------------------------
def ref_stage1(L, model, left, right, min_disp, max_disp, a_p=0.0, a_sm=0.2 * 2 / 512):
    """Train_Stage1_K.py:236-258.  L = the reference's loss_functions module."""
    W = left.shape[3]
    rpan, ldisp = model(left, min_disp, max_disp, ret_disp=True, ret_pan=True, ret_subocc=False)
    vgg_right = L.vgg(right) if a_p > 0 else None
    rec = L.rec_loss_fnc(1, rpan, right, vgg_right, a_p)
    sm = 0
    if a_sm > 0:
        sm = L.smoothness(left[:, :, :, int(0.20 * W)::], ldisp[:, :, :, int(0.20 * W)::], gamma=2)
    return rec + a_sm * sm, rec, sm, rpan, ldisp


def ref_stage2(L, model, fix_model, left, right, min_disp, max_disp, a_p=0.01, a_sm=0.4 * 2 / 512, a_mr=1.0,
               flip=None):
    """Train_Stage2_K.py:247-327.  ``flip``: None = the reference's bilinear grid_sample flip; tests pass an exact
    index flip to both sides (the reference's own flip is only ~1e-4 exact)."""
    import torch
    import torch.nn.functional as F
    B, C, H, W = left.shape
    if flip is None:
        th = torch.zeros(B, 2, 3, device=left.device)
        th[:, 0, 0] = 1
        th[:, 1, 1] = 1
        fg = F.affine_grid(th, [B, C, H, W], align_corners=True).clone()
        fg[:, :, :, 0] = -fg[:, :, :, 0]
        flip = lambda t: F.grid_sample(t, fg, align_corners=True)
    mn2, mx2 = torch.cat((min_disp, min_disp), 0), torch.cat((max_disp, max_disp), 0)
    if a_mr > 0:
        with torch.no_grad():
            d = fix_model(torch.cat((flip(left), right), 0), mn2, mx2, ret_disp=True, ret_pan=False, ret_subocc=False)
            mldisp = flip(d[0:B]).detach()
            mrdisp = d[B:].detach()
    pan, disp, mask0, mask1 = model(torch.cat((left, flip(right)), 0), mn2, mx2,
                                    ret_disp=True, ret_pan=True, ret_subocc=True)
    rpan, lpan = pan[0:B], flip(pan[B:])
    ldisp, rdisp = disp[0:B], flip(disp[B:])
    lmask, rmask = mask0[0:B], flip(mask0[B:])
    rlmask, lrmask = mask1[0:B], flip(mask1[B:])
    vgg_right = L.vgg(right) if a_p > 0 else None
    vgg_left = L.vgg(left) if a_p > 0 else None
    c20, c80 = int(0.20 * W), int(0.80 * W)
    O_L = lmask * lrmask
    O_L[:, :, :, 0:c20] = 1
    O_R = rmask * rlmask
    O_R[:, :, :, c80::] = 1
    if a_mr == 0:
        O_L = O_R = 1
    rec = (L.rec_loss_fnc(O_R, rpan, right, vgg_right, a_p) + L.rec_loss_fnc(O_L, lpan, left, vgg_left, a_p)) / 2
    sm = 0
    if a_sm > 0:
        sm = (L.smoothness(left[:, :, :, c20::], ldisp[:, :, :, c20::], gamma=2) +
              L.smoothness(right[:, :, :, 0:c80], rdisp[:, :, :, 0:c80], gamma=2)) / 2
    mirror = 0
    if a_mr > 0:
        nmaxl = 1 / F.max_pool2d(mldisp, kernel_size=(H, W))
        nmaxr = 1 / F.max_pool2d(mrdisp, kernel_size=(H, W))
        mirror = (torch.mean(nmaxl * (1 - O_L)[:, :, :, c20::] * torch.abs(ldisp - mldisp)[:, :, :, c20::]) +
                  torch.mean(nmaxr * (1 - O_R)[:, :, :, 0:c80] * torch.abs(rdisp - mrdisp)[:, :, :, 0:c80])) / 2
    loss = rec + a_sm * sm + a_mr * mirror
    return dict(loss=loss, rec=rec, sm=sm, mirror=mirror, rpan=rpan, lpan=lpan, ldisp=ldisp, rdisp=rdisp,
                O_L=O_L, O_R=O_R)
