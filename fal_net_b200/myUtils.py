"""Validation metrics of the hot path on the device (csrc/postproc.cu), under the reference's names.

Mirrors the part of /root/reference/myUtils.py that ``validate()`` / ``Test_KITTI`` call per batch
(``get_rmse`` :138-150, ``compute_kitti_errors`` :196-232, ``disps_to_depths_kitti2015`` :234-254,
``disps_to_depths_kitti`` :256-277, the meters :59-110).  The reference moves every prediction to the host
(``.cpu().numpy()``, Train_Stage1_K.py:318-320, Test_KITTI.py:258-266) and evaluates in numpy; here the depth conversion,
the Eigen crop and the seven error sums are ONE kernel per batch, results stay device tensors, and the meters accumulate
on the device (one host read per epoch, when they are printed).  Checkpoint / colour-map / PLY helpers are out of scope
(SURVEY.md 2.1).  No CPU path.
"""
from __future__ import annotations

import torch

from . import _lib

kitti_error_names = ['abs_rel', 'sq_rel', 'rms', 'log_rms', 'a1', 'a2', 'a3']

width_to_focal = {1242: 721.5377, 1241: 718.856, 1224: 707.0493, 1238: 718.3351, 1226: 707.0912, 1280: 738.2355}
width_to_baseline = {1242: 0.9982 * 0.54, 1241: 0.9848 * 0.54, 1224: 1.0144 * 0.54, 1238: 0.9847 * 0.54,
                     1226: 0.9765 * 0.54, 1280: 0.54}


def get_rmse(output_right, label_right, mean=(0.411, 0.432, 0.45)):
    """sqrt(mean((clamp((o + mean) * 255, 0, 255) - (l + mean) * 255)^2)) as a 0-d device tensor (:138-150)."""
    o, l = _lib.f32c(output_right, "output_right"), _lib.f32c(label_right, "label_right")
    B, C, H, W = o.shape
    assert C == 3 and l.shape == o.shape
    s = torch.empty(1, device=o.device, dtype=torch.float64)
    _lib.check(_lib.lib().faln_rmse255(_lib.ptr(o), _lib.ptr(l), _lib.ptr(s), B, H, W, float(mean[0]), float(mean[1]),
                                       float(mean[2]), _lib.cur_stream()), "faln_rmse255")
    return torch.sqrt(s[0] / o.numel()).float()


class _DepthSpec:
    """What disps_to_depths_* hands to compute_kitti_errors: the raw device maps plus how to read them as depths."""

    def __init__(self, t, mode, fb, window):
        self.t, self.mode, self.fb, self.window = t, mode, fb, window


def _maps(t):
    t = _lib.f32c(t)
    if t.dim() == 4:
        t = t[:, 0]
    if t.dim() == 2:
        t = t[None]
    return t.contiguous()


def disps_to_depths_kitti2015(gt_disparities, pred_disparities):
    """:234-254: both sides depth = focal(width) * 0.54 / disparity (gt masked by gt > 0).  Takes [B,H,W] (or
    [B,1,H,W]) device tensors; returns per-image lists like the reference (lazy: the division happens in the kernel)."""
    gt, pr = _maps(gt_disparities), _maps(pred_disparities)
    H, W = gt.shape[1:]
    fb = width_to_focal[W] * 0.54
    win = (0, H, 0, W)
    return ([_DepthSpec(gt[i:i + 1], 0, fb, win) for i in range(gt.shape[0])],
            [_DepthSpec(pr[i:i + 1], 0, fb, win) for i in range(pr.shape[0])])


def disps_to_depths_kitti(gt_disparities, pred_disparities):
    """:256-277 (Eigen split): gt is a depth map; crop [H-219:H-4, 44:1180]; pred depth = focal * baseline / disparity."""
    gt, pr = _maps(gt_disparities), _maps(pred_disparities)
    H, W = gt.shape[1:]
    fb = width_to_focal[W] * width_to_baseline[W]
    win = (H - 219, H - 4, 44, 1180)
    return ([_DepthSpec(gt[i:i + 1], 1, fb, win) for i in range(gt.shape[0])],
            [_DepthSpec(pr[i:i + 1], 1, fb, win) for i in range(pr.shape[0])])


def kitti_error_sums(gt, pred, mode, fb_gt, fb_pred, window, min_d=1.0, max_d=80.0):
    """[B,8] fp64 {count, abs_rel, sq_rel, sq, log_sq, n_a1, n_a2, n_a3} sums per image: one launch for the batch."""
    gt, pred = _maps(gt), _maps(pred)
    B, H, W = gt.shape
    assert pred.shape == gt.shape
    sums = torch.empty(B, 8, device=gt.device, dtype=torch.float64)
    y0, y1, x0, x1 = window
    _lib.check(_lib.lib().faln_kitti_errors(_lib.ptr(gt), _lib.ptr(pred), _lib.ptr(sums), B, H, W, y0, y1, x0, x1, int(mode),
                                            float(fb_gt), float(fb_pred), float(min_d), float(max_d), _lib.cur_stream()),
               "faln_kitti_errors")
    return sums


def errors_from_sums(sums):
    """[...,8] sums -> [...,7] (abs_rel, sq_rel, rms, log_rms, a1, a2, a3), on the device."""
    n = sums[..., 0:1]
    m = sums[..., 1:] / n
    return torch.cat((m[..., 0:2], torch.sqrt(m[..., 2:4]), m[..., 4:7]), dim=-1)


def compute_kitti_errors(gt, pred, use_median=False, min_d=1.0, max_d=80.0):
    """:196-232.  ``gt`` / ``pred`` are the objects returned by disps_to_depths_* (one image each); returns a [7] fp64
    device tensor in the order of ``kitti_error_names``."""
    if use_median:
        raise NotImplementedError("median scaling (stereo-trained FAL-net does not use it: reference default False)")
    assert isinstance(gt, _DepthSpec) and isinstance(pred, _DepthSpec) and gt.mode == pred.mode
    sums = kitti_error_sums(gt.t, pred.t, gt.mode, gt.fb, pred.fb, gt.window, min_d, max_d)
    return errors_from_sums(sums)[0]


def kitti_errors_batch(target, disp, dataset="Kitti2015", min_d=1.0, max_d=80.0):
    """Whole batch in one launch: [B,7] device tensor.  ``dataset``: 'Kitti2015' (disparity ground truth) or
    'Kitti_eigen_test_improved' / 'eigen' (depth ground truth, Eigen crop)."""
    gt, pr = _maps(target), _maps(disp)
    H, W = gt.shape[1:]
    if dataset == "Kitti2015":
        fb = width_to_focal[W] * 0.54
        sums = kitti_error_sums(gt, pr, 0, fb, fb, (0, H, 0, W), min_d, max_d)
    else:
        fb = width_to_focal[W] * width_to_baseline[W]
        sums = kitti_error_sums(gt, pr, 1, fb, fb, (H - 219, H - 4, 44, 1180), min_d, max_d)
    return errors_from_sums(sums)


class AverageMeter(object):
    """:59-78.  ``val`` may be a device tensor: the running sum then lives on the device and nothing synchronises until
    the meter is read (``float(meter.avg)`` / print)."""

    def __init__(self):
        self.reset()

    def reset(self):
        self.val = 0
        self.avg = 0
        self.sum = 0
        self.count = 0

    def update(self, val, n=1):
        if torch.is_tensor(val):
            val = val.detach()
        self.val = val
        self.sum = self.sum + val * n
        self.count += n
        self.avg = self.sum / self.count

    def __repr__(self):
        return 'last:{:.3f} avg:({:.3f})'.format(float(self.val), float(self.avg))


class multiAverageMeter(object):
    """:81-110, vector-valued; accepts a device tensor of len(labels) values."""

    def __init__(self, labels):
        self.meter_no = len(labels)
        self.labels = labels
        self.reset()

    def reset(self):
        self.val = self.avg = self.sum = None
        self.count = 0

    def update(self, val, n=1):
        val = val.detach().double() if torch.is_tensor(val) else torch.as_tensor(val, dtype=torch.float64)
        self.val = val
        self.sum = val * n if self.sum is None else self.sum + val * n
        self.count += n
        self.avg = self.sum / self.count

    def __repr__(self):
        avg = [float("nan")] * self.meter_no if self.avg is None else [float(v) for v in self.avg.cpu()]
        top = "".join("{:>10}".format(l) for l in self.labels)
        return top + "\n" + "".join("{:10.4f}".format(v) for v in avg)
