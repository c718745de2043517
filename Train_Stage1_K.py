#!/usr/bin/env python
"""Stage-1 training entry point on the B200-native hot path.

Mirrors /root/reference/Train_Stage1_K.py: same flag names (:32-70), ``main() / train() / validate()``, Adam with
betas (momentum, beta) and two parameter groups' worth of parameters (:177-181), MultiStepLR milestones (:182),
checkpoint dict {'epoch','m_model','state_dict','best_rmse'} (:202-207), ``validate()`` with RMSE / realEPE / the KITTI
errors computed on the device.  The dataset FILE loaders are out of scope: batches come from ``--synthetic``
KITTI-shaped tensors (``--gpu-augment``: decoded uint8 pairs through the device input pipeline) unless a loader is plugged
in through ``train(train_loader=...)``.  Launch with torchrun for data parallelism (one process per GPU)."""
import argparse
import os
import time

import torch

from fal_net_b200 import models, steps
from fal_net_b200 import loss_functions as LF
from fal_net_b200 import myUtils as utils
from fal_net_b200.entry_common import (AverageMeter, SyntheticRawStereo, SyntheticStereo, SyntheticValidation,
                                       init_distributed, save_checkpoint)
from fal_net_b200.loss_functions import realEPE
from fal_net_b200.trainer import FlatAdamDDP

parser = argparse.ArgumentParser(description="FAL_net in pytorch (B200-native hot path)",
                                 formatter_class=argparse.ArgumentDefaultsHelpFormatter)
parser.add_argument("-d", "--data", metavar="DIR", default="", help="path to dataset (unused with --synthetic)")
parser.add_argument("-n0", "--dataName0", default="Kitti")
parser.add_argument("-maxd", "--max_disp", type=float, default=300)
parser.add_argument("-mind", "--min_disp", type=float, default=2)
parser.add_argument("-gpu_no", "--gpu_no", default="0")
parser.add_argument("-mm", "--m_model", default="FAL_netB", choices=list(models.__all__))
parser.add_argument("-no_levels", "--no_levels", type=int, default=49)
parser.add_argument("-perc", "--a_p", type=float, default=0.01, help="Perceptual loss weight")
parser.add_argument("-smooth", "--a_sm", type=float, default=0.2 * 2 / 512, help="Smoothness loss weight")
parser.add_argument("-b", "--batch_size", type=int, default=8)
parser.add_argument("-ch", "--crop_height", type=int, default=192)
parser.add_argument("-cw", "--crop_width", type=int, default=640)
parser.add_argument("-op", "--optimizer", default="adam")
parser.add_argument("--lr", type=float, default=0.0001)
parser.add_argument("--beta", type=float, default=0.999)
parser.add_argument("--momentum", type=float, default=0.5)
parser.add_argument("--milestones", default=[30, 40], type=int, nargs="*")
parser.add_argument("--weight-decay", "--wd", type=float, default=0.0)
parser.add_argument("--bias-decay", type=float, default=0.0)
parser.add_argument("--epochs", type=int, default=50)
parser.add_argument("--epoch_size", type=int, default=0)
parser.add_argument("--print-freq", "-p", type=int, default=100)
parser.add_argument("--start-epoch", type=int, default=0)
parser.add_argument("--pretrained", default=None, help="checkpoint to resume from")
parser.add_argument("-relbase_test", "--rel_baset", default=1, type=float, help="Relative baseline of testing dataset")
parser.add_argument("-tbs", "--tbatch_size", default=1, type=int, help="validation batch size")
# accepted for command-line compatibility with the reference (the dataset file loaders are out of scope):
# the values are not used
for _opts, _dflt in ((("-train_split", "--train_split"), "eigen_train_split"), (("-vdn", "--vdataName"), "Kitti2015"),
                     (("-w", "--workers"), 4)):
    parser.add_argument(*_opts, default=_dflt, help="accepted, unused (no dataset files on the box)")
parser.add_argument("--sparse", action="store_true", default=True, help="accepted, unused")
parser.add_argument("--synthetic", type=int, default=50, help="synthetic batches per epoch (no dataset on the box)")
parser.add_argument("--val-batches", type=int, default=4, help="synthetic KITTI2015-shaped validation batches per epoch")
parser.add_argument("--gpu-augment", action="store_true",
                    help="feed decoded uint8 375x1242 pairs through the device input pipeline (data_transforms.py on the GPU) "
                         "instead of ready-made crops")
parser.add_argument("--save_path", default="Kitti_stage1")


def train(train_loader, m_model, g_optimizer, epoch, args, device):
    """Step loop of /root/reference/Train_Stage1_K.py:210-276; the loss stays on the device (no per-step .cpu())."""
    batch_time, losses, rec_losses = AverageMeter(), AverageMeter(), AverageMeter()
    epoch_size = len(train_loader) if args.epoch_size == 0 else min(len(train_loader), args.epoch_size)
    m_model.train()
    end = time.time()
    loss_sum, n_steps = None, 0                         # every step's loss, accumulated on the device: one sync per epoch
    for i, ((left_view, right_view), max_disp) in enumerate(train_loader):
        left_view = left_view.to(device, non_blocking=True)
        right_view = right_view.to(device, non_blocking=True)
        max_disp = max_disp.to(device).unsqueeze(1).unsqueeze(1).float()
        min_disp = max_disp * args.min_disp / args.max_disp
        g_optimizer.zero_grad()
        loss, rec_loss, _, _, _ = steps.stage1_loss(m_model, left_view, right_view, min_disp, max_disp, a_p=args.a_p,
                                                    a_sm=args.a_sm)
        loss.backward()
        g_optimizer.step()
        loss_sum = loss.detach() if loss_sum is None else loss_sum + loss.detach()
        n_steps += 1
        if i % args.print_freq == 0:                       # the only host sync, every print_freq steps
            losses.update(loss.item(), args.batch_size)
            rec_losses.update(rec_loss.item(), args.batch_size)
            batch_time.update(time.time() - end)
            print(f"Epoch: [{epoch}][{i}/{epoch_size}] Time {batch_time}  Loss {losses} RecLoss {rec_losses}")
        end = time.time()
        if i >= epoch_size:
            break
    return float(loss_sum) / max(n_steps, 1) if loss_sum is not None else 0.0


@torch.no_grad()
def validate(val_loader, m_model, epoch, args, device):
    """/root/reference/Train_Stage1_K.py:279-347 without a single device-to-host copy inside the loop: RMSE of the synthesised
    view, realEPE and the seven KITTI depth errors are device kernels (fal_net_b200.myUtils / loss_functions.realEPE) and
    the meters accumulate device tensors; the host reads them once, when they are printed."""
    RMSES, EPEs = utils.AverageMeter(), utils.AverageMeter()
    kitti_erros = utils.multiAverageMeter(utils.kitti_error_names)
    m_model.eval()
    for (input_left, input_right), target in val_loader:
        input_left = input_left.to(device, non_blocking=True)
        input_right = input_right.to(device, non_blocking=True)
        target = target.to(device, non_blocking=True)
        B = input_left.shape[0]
        max_disp = torch.full((B, 1, 1), float(args.max_disp) * float(args.rel_baset), device=device)
        min_disp = max_disp * args.min_disp / args.max_disp
        p_im, disp, maskL, maskRL = m_model(input_left, min_disp, max_disp, ret_disp=True, ret_pan=True, ret_subocc=True)
        RMSES.update(utils.get_rmse(p_im, input_right))
        EPEs.update(realEPE(disp, target, sparse=args.sparse), B)
        kitti_erros.update(utils.kitti_errors_batch(target, disp, "Kitti2015").mean(0), B)
    print("* RMSE {0}".format(float(RMSES.avg)))
    print(" * EPE {:.3f}".format(float(EPEs.avg)))
    print(kitti_erros)
    return float(RMSES.avg)


def _augmented(raw_loader, aug, max_disp, device):
    """Adapter: decoded uint8 pairs -> the ((left, right), max_disp) batches train() consumes, augmented on the device."""
    class _It:
        def __len__(self):
            return len(raw_loader)

        def __iter__(self):
            for lefts, rights in raw_loader:
                l, r, _ = aug([t.to(device, non_blocking=True) for t in lefts], [t.to(device, non_blocking=True) for t in rights])
                yield (l, r), torch.full((l.shape[0],), float(max_disp))
    return _It()


def main(argv=None):
    args = parser.parse_args(argv)
    rank, world, device = init_distributed()
    network_data = torch.load(args.pretrained, map_location="cpu") if args.pretrained else None
    if network_data:
        args.m_model = network_data["m_model"]
    m_model = models.__dict__[args.m_model](network_data, no_levels=args.no_levels).to(device)
    g_optimizer = FlatAdamDDP(m_model, lr=args.lr, betas=(args.momentum, args.beta), weight_decay=args.weight_decay,
                              bias_decay=args.bias_decay)
    g_optimizer.broadcast_parameters()
    lr = args.lr
    if args.gpu_augment:
        from fal_net_b200.input_pipeline import GpuStereoAugment
        loader = _augmented(SyntheticRawStereo(args.synthetic, args.batch_size, seed=rank),
                            GpuStereoAugment((args.crop_height, args.crop_width), max_pix=args.max_disp), args.max_disp, device)
    else:
        loader = SyntheticStereo(args.synthetic, args.batch_size, args.crop_height, args.crop_width, args.max_disp, seed=rank)
    val_loader = SyntheticValidation(args.val_batches, int(args.tbatch_size))
    best_rmse = -1
    for epoch in range(args.start_epoch, args.epochs):
        lr_e = args.lr * (0.5 ** sum(epoch >= m for m in args.milestones))       # MultiStepLR(gamma=0.5), :182
        if lr_e != lr:
            lr = lr_e
            g_optimizer.set_lr(lr)
        train_loss = train(loader, m_model, g_optimizer, epoch, args, device)
        rmse = validate(val_loader, m_model, epoch, args, device) if args.val_batches > 0 else -1
        is_best = best_rmse < 0 or rmse < best_rmse                                  # reference :195-199
        best_rmse = rmse if is_best else best_rmse
        if rank == 0:
            save_checkpoint({"epoch": epoch + 1, "m_model": args.m_model, "state_dict": m_model.state_dict(),
                             "best_rmse": best_rmse}, is_best, args.save_path)
            print(f"epoch {epoch}: train loss {train_loss:.5f}  val RMSE {rmse:.4f}")


if __name__ == "__main__":
    main()
