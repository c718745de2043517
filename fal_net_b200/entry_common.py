"""Shared pieces of the three entry points: synthetic KITTI-shaped batches (the reference's data loaders --
Datasets/*.py, data_transforms.py -- are out of scope; only their output contract is kept:
float32 NCHW, img/255 - [0.411, 0.432, 0.45], Train_Stage1_K.py:124-128), meters, checkpoint dict."""
from __future__ import annotations

import os
import time

import torch

MEAN = (0.411, 0.432, 0.45)


class SyntheticStereo:
    """Iterable of ((left, right), max_disp) batches shaped like the reference's train loader output
    (Datasets/listdataset_train.py:98: inputs, x_pix)."""

    def __init__(self, n_batches, batch, H, W, max_disp=300.0, seed=0, device="cpu"):
        self.n, self.B, self.H, self.W, self.max_disp, self.seed, self.device = n_batches, batch, H, W, max_disp, seed, device

    def __len__(self):
        return self.n

    def __iter__(self):
        g = torch.Generator().manual_seed(self.seed)
        mean = torch.tensor(MEAN).view(1, 3, 1, 1)
        for _ in range(self.n):
            left = torch.rand(self.B, 3, self.H, self.W, generator=g) - mean
            right = torch.rand(self.B, 3, self.H, self.W, generator=g) - mean
            yield (left.pin_memory() if torch.cuda.is_available() else left,
                   right.pin_memory() if torch.cuda.is_available() else right), torch.full((self.B,), float(self.max_disp))


class SyntheticValidation:
    """Iterable of ((left, right), sparse disparity target) batches shaped like the reference's KITTI2015 validation loader
    output (Datasets/Kitti2015.py; Train_Stage1_K.py:293-297): 375x1242 views and a ~25 %-dense disparity map, zeros =
    invalid."""

    def __init__(self, n_batches, batch=1, H=375, W=1242, seed=7):
        self.n, self.B, self.H, self.W, self.seed = n_batches, batch, H, W, seed

    def __len__(self):
        return self.n

    def __iter__(self):
        g = torch.Generator().manual_seed(self.seed)
        mean = torch.tensor(MEAN).view(1, 3, 1, 1)
        for _ in range(self.n):
            left = torch.rand(self.B, 3, self.H, self.W, generator=g) - mean
            right = torch.rand(self.B, 3, self.H, self.W, generator=g) - mean
            disp = 1.0 + 150.0 * torch.rand(self.B, 1, self.H, self.W, generator=g)
            keep = torch.rand(self.B, 1, self.H, self.W, generator=g) < 0.25
            yield (left, right), disp * keep


class SyntheticRawStereo:
    """Decoded uint8 stereo pairs [H,W,3] as the reference's image loader hands them to the co-transforms
    (Datasets/listdataset_train.py:77-80), for the device input pipeline (fal_net_b200.input_pipeline)."""

    def __init__(self, n_batches, batch, H=375, W=1242, seed=0):
        self.n, self.B, self.H, self.W, self.seed = n_batches, batch, H, W, seed

    def __len__(self):
        return self.n

    def __iter__(self):
        g = torch.Generator().manual_seed(self.seed)
        for _ in range(self.n):
            yield ([torch.randint(0, 256, (self.H, self.W, 3), generator=g, dtype=torch.uint8) for _ in range(self.B)],
                   [torch.randint(0, 256, (self.H, self.W, 3), generator=g, dtype=torch.uint8) for _ in range(self.B)])


class AverageMeter:
    def __init__(self):
        self.sum, self.count, self.val = 0.0, 0, 0.0

    def update(self, val, n=1):
        self.val = float(val)
        self.sum += float(val) * n
        self.count += n

    @property
    def avg(self):
        return self.sum / max(self.count, 1)

    def __repr__(self):
        return f"{self.val:.4f} ({self.avg:.4f})"


def save_checkpoint(state, is_best, save_path, filename="checkpoint.pth.tar"):
    """Same dict and file names as /root/reference/myUtils.py:10-13."""
    os.makedirs(save_path, exist_ok=True)
    torch.save(state, os.path.join(save_path, filename))
    if is_best:
        import shutil
        shutil.copyfile(os.path.join(save_path, filename), os.path.join(save_path, "model_best.pth.tar"))


def init_distributed():
    """One process per GPU when launched by torchrun; single process otherwise."""
    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1 and not dist.is_initialized():
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    return int(os.environ.get("RANK", "0")), world, torch.device("cuda", local)


now = time.time
