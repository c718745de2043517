"""Step bodies of the three entry points, on the fused kernels.

  stage1_loss  <- /root/reference/Train_Stage1_K.py:236-258
  stage2_loss  <- /root/reference/Train_Stage2_K.py:247-327
  stage1_slow_loss <- /root/reference/Train_Stage1_Kslow.py:236-278
  test_disp    <- /root/reference/Test_KITTI.py:196-205, 287-300

Differences from the reference that do not change results beyond fp rounding:
  * horizontal flips are exact index reversals folded into kernel indexing (``flip_x``) instead of seven
    bilinear grid_sample calls per step (the reference's flip is only ~1e-4 exact);
  * no ``.cpu()`` synchronisation inside the step: losses stay on the device.
"""
from __future__ import annotations

import torch

import os

from . import loss_functions as LF
from . import losses as K

_OVERLAP = os.environ.get("FALN_STAGE2_OVERLAP", "1") not in ("", "0")
_AUX_STREAMS: dict = {}


def _aux_stream(device):
    key = torch.device(device).index if torch.device(device).index is not None else torch.cuda.current_device()
    st = _AUX_STREAMS.get(key)
    if st is None:
        st = _AUX_STREAMS[key] = torch.cuda.Stream(device=key)
    return st


def stage1_loss(model, left, right, min_disp, max_disp, a_p=0.01, a_sm=0.2 * 2 / 512, vgg=None):
    """Returns (loss, rec_loss, sm_loss, rpan, ldisp); all device tensors."""
    W = left.shape[3]
    rpan, ldisp = model(left, min_disp, max_disp, ret_disp=True, ret_pan=True, ret_subocc=False)
    vgg_right = None
    if a_p > 0:
        with torch.no_grad():
            vgg_right = (vgg or LF.vgg)(right)
    rec = LF.rec_loss_fnc(1, rpan, right, vgg_right, a_p)
    sm = 0
    if a_sm > 0:
        # the 20 % left dis-occluded band has no supervision (Train_Stage1_K.py:254-255)
        sm = LF.smoothness(left, ldisp, gamma=2, window=(int(0.20 * W), W))
    loss = LF.combine([(1.0, rec), (a_sm, sm)])
    return loss, rec, sm, rpan, ldisp


def stage2_loss(model, fix_model, left, right, min_disp, max_disp, a_p=0.01, a_sm=0.4 * 2 / 512, a_mr=1.0, vgg=None):
    """Returns dict(loss, rec, sm, mirror, ...).  ``fix_model`` is the frozen Stage-1 network."""
    B, _, H, W = left.shape
    c20, c80 = int(0.20 * W), int(0.80 * W)
    mn2, mx2 = torch.cat((min_disp, min_disp), 0), torch.cat((max_disp, max_disp), 0)
    left_f, right_f = torch.flip(left, dims=[3]), torch.flip(right, dims=[3])
    vgg = vgg or LF.vgg

    # The frozen model's pass and the VGG features of the two label views depend on the inputs only: they run on a side
    # stream beside the trainable forward (a parallel branch of the captured step graph), so that one chain's large layers
    # fill the SMs the other chain's small-map layers leave idle.  FALN_STAGE2_OVERLAP=0 runs them in line.
    mldisp = mrdisp = vgg_right = vgg_left = None
    side = _aux_stream(left.device) if _OVERLAP else None
    main = torch.cuda.current_stream(left.device)

    def frozen_work():
        nonlocal mldisp, mrdisp, vgg_right, vgg_left
        with torch.no_grad():
            if a_mr > 0:                                                 # :255-264
                dfix = fix_model(torch.cat((left_f, right), 0), mn2, mx2, ret_disp=True, ret_pan=False, ret_subocc=False)
                mldisp = torch.flip(dfix[:B], dims=[3]).contiguous()
                mrdisp = dfix[B:].contiguous()
            if a_p > 0:
                vgg_right, vgg_left = vgg(right), vgg(left)

    if side is not None:
        side.wait_stream(main)
        with torch.cuda.stream(side):
            frozen_work()
    else:
        frozen_work()

    pan, disp, mask0, mask1 = model(torch.cat((left, right_f), 0), mn2, mx2,
                                    ret_disp=True, ret_pan=True, ret_subocc=True)        # :267-271
    # second half of the batch lives in flipped coordinates; the un-flip (:283-286) is folded into the kernels
    rpan, lpan_f = pan[:B], pan[B:]
    ldisp, rdisp_f = disp[:B], disp[B:]

    if side is not None:
        main.wait_stream(side)
        for t in (mldisp, mrdisp) + tuple(vgg_right or ()) + tuple(vgg_left or ()):
            if t is not None:
                t.record_stream(main)                                    # allocated on the side stream, consumed here

    if a_mr > 0:                                                         # :295-299
        O_L = K.occ_mask(mask0[:B], mask1[B:], False, True, 0, c20)      # lmask * unflip(lrmask); first 20 % := 1
        O_R = K.occ_mask(mask0[B:], mask1[:B], True, False, c80, W)      # unflip(rmask) * rlmask; last 20 % := 1
    else:
        O_L = O_R = 1

    rec = LF.combine([(0.5, LF.rec_loss_fnc(O_R, rpan, right, vgg_right, a_p)),
                      (0.5, LF.rec_loss_fnc(O_L, lpan_f, left, vgg_left, a_p, flip_x=True))])
    sm = 0
    if a_sm > 0:
        sm = LF.combine([(0.5, LF.smoothness(left, ldisp, gamma=2, window=(c20, W))),
                         (0.5, LF.smoothness(right, rdisp_f, gamma=2, window=(0, c80), flip_x=True))])
    mirror = 0
    if a_mr > 0:
        mirror = LF.combine([(0.5, LF.mirror_loss(ldisp, mldisp, O_L, (c20, W))),
                             (0.5, LF.mirror_loss(rdisp_f, mrdisp, O_R, (0, c80), flip_x=True))])
    loss = LF.combine([(1.0, rec), (a_sm, sm), (a_mr, mirror)])
    return dict(loss=loss, rec=rec, sm=sm, mirror=mirror, rpan=rpan, lpan_f=lpan_f, ldisp=ldisp, rdisp_f=rdisp_f,
                O_L=O_L, O_R=O_R)


def stage1_slow_loss(model, left, right, min_disp, max_disp, a_p=0.01, a_sm=0.2 * 2 / 512, vgg=None):
    """Step body of /root/reference/Train_Stage1_Kslow.py:236-278: both views through the network as one batch
    ([left, flip(right)]), reconstruction + smoothness on both, no occlusion masks, no mirror loss.  Returns dict."""
    B, _, H, W = left.shape
    c20, c80 = int(0.20 * W), int(0.80 * W)
    mn2, mx2 = torch.cat((min_disp, min_disp), 0), torch.cat((max_disp, max_disp), 0)
    vgg = vgg or LF.vgg
    pan, disp = model(torch.cat((left, torch.flip(right, dims=[3])), 0), mn2, mx2, ret_disp=True, ret_pan=True,
                      ret_subocc=False)
    rpan, lpan_f = pan[:B], pan[B:]                      # second half lives in flipped coordinates (un-flip folded below)
    ldisp, rdisp_f = disp[:B], disp[B:]
    vgg_right = vgg_left = None
    if a_p > 0:
        with torch.no_grad():
            vgg_right, vgg_left = vgg(right), vgg(left)
    rec = LF.combine([(0.5, LF.rec_loss_fnc(1, rpan, right, vgg_right, a_p)),
                      (0.5, LF.rec_loss_fnc(1, lpan_f, left, vgg_left, a_p, flip_x=True))])
    sm = 0
    if a_sm > 0:
        sm = LF.combine([(0.5, LF.smoothness(left, ldisp, gamma=2, window=(c20, W))),
                         (0.5, LF.smoothness(right, rdisp_f, gamma=2, window=(0, c80), flip_x=True))])
    loss = LF.combine([(1.0, rec), (a_sm, sm)])
    return dict(loss=loss, rec=rec, sm=sm, rpan=rpan, lpan_f=lpan_f, ldisp=ldisp, rdisp_f=rdisp_f)


@torch.no_grad()
def test_disp(model, image, min_disp, max_disp, f_post_process=False, ms_post_process=False):
    """Disparity prediction with the reference's two post-processing modes (Test_KITTI.py:196-205)."""
    disp = model(image, min_disp, max_disp, ret_disp=True, ret_subocc=False, ret_pan=False)
    if f_post_process:
        flip_disp = model(torch.flip(image, dims=[3]), min_disp, max_disp, ret_disp=True, ret_subocc=False, ret_pan=False)
        disp = (disp + torch.flip(flip_disp, dims=[3])) / 2
    elif ms_post_process:
        disp = ms_pp(image, model, disp, min_disp, max_disp)
    return disp


@torch.no_grad()
def ms_pp(input_view, model, disp, min_disp, max_pix):
    """Multi-scale post-processing (Test_KITTI.py:287-300) on three small kernels (csrc/postproc.cu) around the 2/3-scale
    network pass: flip + bilinear down-scale of the view; an exact device-side 95th percentile PER IMAGE (the reference
    runs batch 1, so its np.percentile is per image; here B > 1 keeps that meaning and nothing syncs with the host);
    nearest up-sampling + un-flip + blend in one pass."""
    from . import postproc
    up_fac = 2 / 3
    small = postproc.flip_resize_bilinear(input_view, scale_factor=up_fac, flip_x=True)
    d2 = model(small, min_disp, max_pix, ret_disp=True, ret_pan=False, ret_subocc=False)
    p95 = postproc.percentile_rows(disp, 95.0, add=1e-6)
    return postproc.mspp_blend(disp, d2, p95, 1 / up_fac)
