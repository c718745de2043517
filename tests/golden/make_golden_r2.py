#!/usr/bin/env python
"""Round-2 golden fixtures, generated FROM THE REFERENCE ITSELF (build container only; CPU):

    python tests/golden/make_golden_r2.py [metrics] [variants] [pipeline]

  metrics.npz    validation metrics of /root/reference/myUtils.py (get_rmse, disps_to_depths_kitti2015 / _kitti +
                 compute_kitti_errors) and loss_functions.realEPE on seeded synthetic sparse ground truth
  variants.npz   FAL_netA / FAL_netC: constructor checksums + forward outputs at a small size (section `variants`)
  pipeline.npz   the training input pipeline (data_transforms.py co-transforms + Normalize) on a seeded uint8 stereo pair

Same shims as make_golden.py (``.cuda()`` = identity, seeded VGG stand-in).  Each section also asserts that the oracle's
restatement reproduces the reference, so the committed vectors pin the oracle as well as the CUDA path.
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)

from make_golden import import_reference, images, disp_range  # noqa: E402
from oracle import falnet_oracle as O  # noqa: E402


def sparse_gt(B, H, W, seed, lo, hi, keep=0.25):
    g = torch.Generator().manual_seed(seed)
    v = lo + (hi - lo) * torch.rand(B, 1, H, W, generator=g)
    m = torch.rand(B, 1, H, W, generator=g) < keep
    return (v * m).float()


def metrics(ref_models, ref_losses):
    sys.path.insert(0, "/root/reference")
    import myUtils as RU
    out = {}
    # ---- KITTI2015-style: disparity ground truth, 375x1242, B=2
    B, H, W = 2, 375, 1242
    gt = sparse_gt(B, H, W, 11, 1.0, 180.0)
    g = torch.Generator().manual_seed(12)
    pred = (gt * (0.8 + 0.4 * torch.rand(gt.shape, generator=g)) + (gt == 0) * 200 * torch.rand(gt.shape, generator=g)).float()
    pred[:, :, :5, :7] = 0.0                                   # some non-positive predictions (pred_mask path)
    td, pd = RU.disps_to_depths_kitti2015(gt.squeeze(1).numpy(), pred.squeeze(1).numpy())
    errs = np.array([RU.compute_kitti_errors(td[i].copy(), pd[i].copy()) for i in range(B)], dtype=np.float64)
    for i in range(B):
        a, b = O.depths_kitti2015(gt[i, 0].numpy(), pred[i, 0].numpy())
        assert np.allclose(O.kitti_errors(a, b), errs[i], rtol=1e-12, atol=0), i
    out["k15_errs"] = errs
    out["k15_meta"] = np.array([B, H, W, 11, 12])
    epe = ref_losses.realEPE(pred, gt, sparse=True)
    assert torch.equal(epe, O.real_epe(pred, gt, sparse=True))
    out["k15_epe_sparse"] = epe.numpy()
    out["k15_epe_dense"] = ref_losses.realEPE(pred, gt, sparse=False).numpy()
    small = torch.nn.functional.avg_pool2d(pred, 3)
    out["k15_epe_upsampled"] = ref_losses.realEPE(small, gt, sparse=True).numpy()
    assert torch.equal(ref_losses.realEPE(small, gt, sparse=True), O.real_epe(small, gt, sparse=True))
    # ---- Eigen-style: depth ground truth with crop
    gt_d = sparse_gt(B, H, W, 13, 0.5, 95.0)
    td, pd = RU.disps_to_depths_kitti(gt_d.squeeze(1).numpy(), pred.squeeze(1).numpy())
    errs = np.array([RU.compute_kitti_errors(td[i].copy(), pd[i].copy()) for i in range(B)], dtype=np.float64)
    for i in range(B):
        a, b = O.depths_kitti_eigen(gt_d[i, 0].numpy(), pred[i, 0].numpy())
        assert np.allclose(O.kitti_errors(a, b), errs[i], rtol=1e-12, atol=0), i
    out["eig_errs"] = errs
    out["eig_meta"] = np.array([B, H, W, 13, 12])
    # ---- RMSE of a synthesised view
    o, l = images(2, 64, 200, 5) * 1.3, images(2, 64, 200, 6)
    r = RU.get_rmse(o, l)
    assert torch.equal(r, O.get_rmse(o, l))
    out["rmse"] = r.numpy()
    np.savez_compressed(os.path.join(HERE, "metrics.npz"), **out)
    print("metrics.npz written:", {k: v.shape for k, v in out.items()})


def main():
    ref_models, ref_losses, _ = import_reference()
    torch.set_num_threads(max(1, os.cpu_count() or 1))
    want = sys.argv[1:] or ["metrics", "variants", "pipeline"]
    if "metrics" in want:
        metrics(ref_models, ref_losses)
    if "variants" in want:
        import make_golden_variants
        make_golden_variants.run(ref_models)
    if "pipeline" in want:
        import make_golden_pipeline
        make_golden_pipeline.run()


if __name__ == "__main__":
    main()
