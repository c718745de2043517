"""Validation metrics and ms_pp pieces.

CPU: the oracle's numpy restatement against the golden values the REFERENCE's myUtils.py / loss_functions.realEPE produced
(tests/golden/metrics.npz, generator tests/golden/make_golden_r2.py).  GPU: the device kernels (csrc/postproc.cu, through
the C ABI) against the same golden values and against the oracle on further seeded cases."""
import numpy as np
import pytest
import torch

from oracle import falnet_oracle as O
from tests.helpers import images, metrics_inputs, rel_err


@pytest.fixture(scope="module")
def gold(golden_dir):
    return np.load(f"{golden_dir}/metrics.npz")


def test_oracle_metrics_match_reference_golden(gold):
    gt, gt_d, pred = metrics_inputs()
    for i in range(gt.shape[0]):
        a, b = O.depths_kitti2015(gt[i, 0].numpy(), pred[i, 0].numpy())
        assert np.allclose(O.kitti_errors(a, b), gold["k15_errs"][i], rtol=1e-12, atol=0)
        a, b = O.depths_kitti_eigen(gt_d[i, 0].numpy(), pred[i, 0].numpy())
        assert np.allclose(O.kitti_errors(a, b), gold["eig_errs"][i], rtol=1e-12, atol=0)
    assert float(O.real_epe(pred, gt, sparse=True)) == float(gold["k15_epe_sparse"])
    assert float(O.real_epe(pred, gt, sparse=False)) == float(gold["k15_epe_dense"])
    small = torch.nn.functional.avg_pool2d(pred, 3)
    assert float(O.real_epe(small, gt, sparse=True)) == float(gold["k15_epe_upsampled"])
    assert float(O.get_rmse(images(2, 64, 200, 5) * 1.3, images(2, 64, 200, 6))) == float(gold["rmse"])


@pytest.mark.gpu
def test_device_kitti_errors_match_reference_golden(gold):
    from fal_net_b200 import myUtils as U
    dev = torch.device("cuda:0")
    gt, gt_d, pred = metrics_inputs()
    e15 = U.kitti_errors_batch(gt.to(dev), pred.to(dev), "Kitti2015").cpu().numpy()
    assert np.allclose(e15, gold["k15_errs"], rtol=1e-9, atol=0), (e15, gold["k15_errs"])
    eig = U.kitti_errors_batch(gt_d.to(dev), pred.to(dev), "eigen").cpu().numpy()
    # the reference takes log() of its float32 ground-truth depths in float32 (myUtils.py:223); the kernel works in fp64
    assert np.allclose(eig, gold["eig_errs"], rtol=2e-7, atol=0), (eig, gold["eig_errs"])
    # the reference's two-step API, image by image
    td, pd = U.disps_to_depths_kitti2015(gt.squeeze(1).to(dev), pred.squeeze(1).to(dev))
    one = U.compute_kitti_errors(td[1], pd[1]).cpu().numpy()
    assert np.allclose(one, gold["k15_errs"][1], rtol=1e-9, atol=0)
    td, pd = U.disps_to_depths_kitti(gt_d.squeeze(1).to(dev), pred.squeeze(1).to(dev))
    assert np.allclose(U.compute_kitti_errors(td[0], pd[0]).cpu().numpy(), gold["eig_errs"][0], rtol=2e-7, atol=0)
    m = U.multiAverageMeter(U.kitti_error_names)
    for i in range(2):
        m.update(torch.from_numpy(gold["k15_errs"][i]).to(dev), 1)
    assert np.allclose(m.avg.cpu().numpy(), gold["k15_errs"].mean(0)) and "abs_rel" in repr(m)


@pytest.mark.gpu
def test_device_epe_and_rmse_match_reference_golden(gold):
    from fal_net_b200 import loss_functions as LF, myUtils as U
    dev = torch.device("cuda:0")
    gt, _, pred = metrics_inputs()
    assert abs(float(LF.realEPE(pred.to(dev), gt.to(dev), sparse=True)) / float(gold["k15_epe_sparse"]) - 1) < 1e-6
    assert abs(float(LF.realEPE(pred.to(dev), gt.to(dev), sparse=False)) / float(gold["k15_epe_dense"]) - 1) < 1e-6
    small = torch.nn.functional.avg_pool2d(pred, 3)
    assert abs(float(LF.realEPE(small.to(dev), gt.to(dev), sparse=True)) / float(gold["k15_epe_upsampled"]) - 1) < 1e-5
    o, l = images(2, 64, 200, 5) * 1.3, images(2, 64, 200, 6)
    assert abs(float(U.get_rmse(o.to(dev), l.to(dev))) / float(gold["rmse"]) - 1) < 1e-6


@pytest.mark.gpu
@pytest.mark.parametrize("B,n", [(1, 465750), (8, 465750), (3, 1001), (2, 7), (4, 1)])
def test_device_percentile_is_numpy_percentile_per_image(B, n):
    """Exact order statistics + numpy's 'linear' interpolation, per row, incl. ties, negatives and tiny rows."""
    from fal_net_b200 import postproc
    dev = torch.device("cuda:0")
    g = torch.Generator().manual_seed(n + B)
    x = (300 * torch.rand(B, n, generator=g)) ** 1.5
    if n > 100:
        x[0, : n // 3] = 17.25                                   # heavy ties
        x[-1] = -x[-1]                                           # negatives
    for q in (95.0, 50.0, 0.0, 100.0, 33.3):
        got = postproc.percentile_rows(x.to(dev), q).cpu().numpy()
        want = np.array([np.percentile(x[b].numpy(), q) for b in range(B)])
        assert np.allclose(got, want.astype(np.float32), rtol=1e-6, atol=0), (q, got, want)


@pytest.mark.gpu
@pytest.mark.parametrize("B,H,W", [(1, 375, 1242), (2, 54, 180), (8, 192, 640)])
def test_device_flip_resize_and_blend_match_aten(B, H, W):
    """The ms_pp kernels against the reference's ATen ops (Test_KITTI.py:291-300) executed on the same GPU."""
    import torch.nn.functional as F
    from fal_net_b200 import postproc
    dev = torch.device("cuda:0")
    img = images(B, H, W, 3).to(dev)
    up_fac = 2 / 3
    want = F.interpolate(torch.flip(img, dims=[3]), scale_factor=up_fac, mode="bilinear", align_corners=True)
    got = postproc.flip_resize_bilinear(img, scale_factor=up_fac)
    assert got.shape == want.shape and rel_err(got, want) < 1e-6
    g = torch.Generator(device=dev).manual_seed(4)
    disp = 120 * torch.rand(B, 1, H, W, generator=g, device=dev)
    small = 80 * torch.rand(B, 1, want.shape[2], want.shape[3], generator=g, device=dev)
    p = postproc.percentile_rows(disp, 95.0, add=1e-6)
    got = postproc.mspp_blend(disp, small, p, 1 / up_fac)
    d2 = torch.flip((1 / up_fac) * F.interpolate(small, size=(H, W), mode="nearest"), dims=[3])
    rows = []
    for b in range(B):                                           # per image, like the batch-1 reference
        norm = disp[b:b + 1] / (np.percentile(disp[b:b + 1].cpu().numpy(), 95) + 1e-6)
        norm[norm > 1] = 1
        rows.append((1 - norm) * disp[b:b + 1] + norm * d2[b:b + 1])
    assert rel_err(got, torch.cat(rows, 0)) < 1e-6


@pytest.mark.gpu
def test_ms_pp_per_image_semantics_batch_vs_single():
    """with B > 1, ms_pp of the batch equals ms_pp of each image alone (the percentile is per image).  The three
    images get disparity ranges 1 : 0.3 : 0.6, so a percentile taken over the batch would move the blend weights of the
    second image by a factor ~3.  Tolerance: the convolution kernels are planned per launch shape (a batch of 3 and a batch
    of 1 use different tilings / K splits, like cuDNN's per-shape algorithm choice), so the two runs are two bf16 evaluations
    of the network: the 2e-2 of BASELINE's north_star, not bit equality."""
    from fal_net_b200 import models, steps
    dev = torch.device("cuda:0")
    torch.manual_seed(0)
    m = models.FAL_netB(no_levels=49).to(dev)
    B, H, W = 3, 96, 320
    img = images(B, H, W, 21).to(dev)
    mx = torch.tensor([300.0, 90.0, 180.0], device=dev).view(B, 1, 1)
    mn = mx * 2 / 300
    whole = steps.test_disp(m, img, mn, mx, ms_post_process=True)
    plain = steps.test_disp(m, img, mn, mx)
    p_all = float(np.percentile(plain.cpu().numpy(), 95))
    for b in range(B):
        one = steps.test_disp(m, img[b:b + 1], mn[b:b + 1], mx[b:b + 1], ms_post_process=True)
        assert rel_err(whole[b:b + 1], one) < 2e-2, b
    # the discriminating power of the check: image 1's own percentile is far from the batch's
    p_1 = float(np.percentile(plain[1:2].cpu().numpy(), 95))
    assert p_1 < 0.6 * p_all, (p_1, p_all)
