#!/usr/bin/env python
"""Stage-1 training entry point on the B200-native hot path.

Mirrors /root/reference/Train_Stage1_K.py: same flag names (:32-70), ``main() / train() / validate()``, Adam with
betas (momentum, beta) and two parameter groups' worth of parameters (:177-181), MultiStepLR milestones (:182),
checkpoint dict {'epoch','m_model','state_dict','best_rmse'} (:202-207).  The dataset loaders are out of scope
(SURVEY.md 2.1): batches come from ``--synthetic`` KITTI-shaped tensors unless a loader is plugged in through
``train(train_loader=...)``.  Launch with torchrun for data parallelism (one process per GPU)."""
import argparse
import os
import time

import torch

from fal_net_b200 import models, steps
from fal_net_b200 import loss_functions as LF
from fal_net_b200.entry_common import AverageMeter, SyntheticStereo, init_distributed, save_checkpoint
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
# accepted for command-line compatibility with the reference (data loading / validation split are out of scope,
# SURVEY.md 2.1): the values are not used
for _opts, _dflt in ((("-train_split", "--train_split"), "eigen_train_split"), (("-vdn", "--vdataName"), "Kitti2015"),
                     (("-relbase_test", "--rel_baset"), 1), (("-w", "--workers"), 4), (("-tbs", "--tbatch_size"), 1)):
    parser.add_argument(*_opts, default=_dflt, help="accepted, unused")
parser.add_argument("--sparse", action="store_true", default=True, help="accepted, unused")
parser.add_argument("--synthetic", type=int, default=50, help="synthetic batches per epoch (no dataset on the box)")
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
    loader = SyntheticStereo(args.synthetic, args.batch_size, args.crop_height, args.crop_width, args.max_disp, seed=rank)
    for epoch in range(args.start_epoch, args.epochs):
        lr_e = args.lr * (0.5 ** sum(epoch >= m for m in args.milestones))       # MultiStepLR(gamma=0.5), :182
        if lr_e != lr:
            lr = lr_e
            g_optimizer.set_lr(lr)
        train_loss = train(loader, m_model, g_optimizer, epoch, args, device)
        if rank == 0:
            save_checkpoint({"epoch": epoch + 1, "m_model": args.m_model, "state_dict": m_model.state_dict(),
                             "best_rmse": -1}, False, args.save_path)
            print(f"epoch {epoch}: train loss {train_loss:.5f}")


if __name__ == "__main__":
    main()
