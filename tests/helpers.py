"""Shared test utilities: seeded synthetic inputs (the same recipes tests/golden/make_golden.py used)."""
import hashlib

import torch

MEAN = (0.411, 0.432, 0.45)


def images(B, H, W, seed):
    g = torch.Generator().manual_seed(seed)
    return torch.rand(B, 3, H, W, generator=g) - torch.tensor(MEAN).view(1, 3, 1, 1)


def disp_range(B, max_disp=300.0, min_disp=2.0):
    mx = torch.full((B, 1, 1), float(max_disp))
    return mx * min_disp / max_disp, mx


def sha(t):
    return hashlib.sha256(t.detach().contiguous().numpy().tobytes()).hexdigest()[:16]


def rel_err(a, b):
    """max|a-b| / max|b|  -- the tolerance metric of SURVEY.md 8c(iv)."""
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))


def med_case_inputs(tag, B, N, H, W):
    """Inputs of golden MED case `tag` (tests/golden/make_golden.py section 2)."""
    g = torch.Generator().manual_seed(7 + len(tag) + N + W)
    logits = 2 * torch.randn(B, N, H, W, generator=g)
    img = images(B, H, W, 1234 + W)
    gp = torch.randn(B, 3, H, W, generator=g)
    gd = torch.randn(B, 1, H, W, generator=g)
    return logits, img, gp, gd


def rel_l2(a, b):
    """||a-b||_2 / ||b||_2."""
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


def sparse_gt(B, H, W, seed, lo, hi, keep=0.25):
    """Seeded sparse ground-truth map (zeros = invalid), as tests/golden/make_golden_r2.py draws it."""
    g = torch.Generator().manual_seed(seed)
    v = lo + (hi - lo) * torch.rand(B, 1, H, W, generator=g)
    m = torch.rand(B, 1, H, W, generator=g) < keep
    return (v * m).float()


def metrics_inputs(B=2, H=375, W=1242):
    """(gt disparity, gt depth, predicted disparity) of the metrics golden cases (make_golden_r2.py `metrics`)."""
    gt = sparse_gt(B, H, W, 11, 1.0, 180.0)
    g = torch.Generator().manual_seed(12)
    pred = (gt * (0.8 + 0.4 * torch.rand(gt.shape, generator=g)) + (gt == 0) * 200 * torch.rand(gt.shape, generator=g)).float()
    pred[:, :, :5, :7] = 0.0
    gt_d = sparse_gt(B, H, W, 13, 0.5, 95.0)
    return gt, gt_d, pred
