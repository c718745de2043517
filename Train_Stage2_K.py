#!/usr/bin/env python
"""Stage-2 training entry point on the B200-native hot path.

Mirrors /root/reference/Train_Stage2_K.py: flag names (:32-71), frozen ``fix_model`` (:176-183), mirrored-occlusion
masks and mirror loss (:247-327), Adam lr 5e-5 (:53).  See Train_Stage1_K.py for what is in / out of scope."""
import time

import torch

import Train_Stage1_K as S1
from fal_net_b200 import models, steps
from fal_net_b200.entry_common import AverageMeter, SyntheticStereo, init_distributed, save_checkpoint
from fal_net_b200.trainer import FlatAdamDDP

parser = S1.parser
# defaults of /root/reference/Train_Stage2_K.py:44-60 where they differ from Stage 1
parser.set_defaults(lr=0.00005, a_sm=0.4 * 2 / 512, batch_size=4, milestones=[5, 10], epochs=20, save_path="Kitti_stage2")
parser.add_argument("-mirror_loss", "--a_mr", type=float, default=1, help="Mirror loss weight")
parser.add_argument("--fix_model", default=None, help="Stage-1 checkpoint of the frozen teacher (reference default: a run "
                    "directory, Train_Stage2_K.py:63-65); required when the mirror loss is on")
parser.add_argument("--random-teacher", action="store_true",
                    help="benchmarking only: allow a random-init frozen model when --fix_model is absent and a_mr > 0")


def train(train_loader, m_model, fix_model, g_optimizer, epoch, args, device):
    """Step loop of /root/reference/Train_Stage2_K.py:220-345."""
    batch_time, losses = AverageMeter(), AverageMeter()
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
        res = steps.stage2_loss(m_model, fix_model, left_view, right_view, min_disp, max_disp, a_p=args.a_p,
                                a_sm=args.a_sm, a_mr=args.a_mr)
        res["loss"].backward()
        g_optimizer.step()
        loss_sum = res["loss"].detach() if loss_sum is None else loss_sum + res["loss"].detach()
        n_steps += 1
        if i % args.print_freq == 0:                    # the only host sync inside the epoch
            losses.update(res["loss"].item(), args.batch_size)
            batch_time.update(time.time() - end)
            print(f"Epoch: [{epoch}][{i}/{epoch_size}] Time {batch_time}  Loss {losses}")
        end = time.time()
        if i >= epoch_size:
            break
    return float(loss_sum) / max(n_steps, 1) if loss_sum is not None else 0.0


def main(argv=None):
    args = parser.parse_args(argv)
    rank, world, device = init_distributed()
    network_data = torch.load(args.pretrained, map_location="cpu") if args.pretrained else None
    if network_data:
        args.m_model = network_data["m_model"]                                   # reference :166
    m_model = models.__dict__[args.m_model](network_data, no_levels=args.no_levels).to(device)
    if args.a_mr > 0 and not args.fix_model and not args.random_teacher:
        # the reference defaults --fix_model to a Stage-1 checkpoint and fails without it (:63-65,176-183); a random frozen
        # teacher would make the mirror loss meaningless, so it has to be asked for explicitly
        raise SystemExit("Train_Stage2_K: --fix_model <stage-1 checkpoint> is required when a_mr > 0 "
                         "(pass --random-teacher for a synthetic benchmark run)")
    fix_data = torch.load(args.fix_model, map_location="cpu") if args.fix_model else None
    fix_name = fix_data["m_model"] if fix_data and "m_model" in fix_data else args.m_model   # reference :178-180
    fix_model = models.__dict__[fix_name](fix_data, no_levels=args.no_levels).to(device).eval()
    for p in fix_model.parameters():
        p.requires_grad_(False)
    g_optimizer = FlatAdamDDP(m_model, lr=args.lr, betas=(args.momentum, args.beta), weight_decay=args.weight_decay,
                              bias_decay=args.bias_decay)
    g_optimizer.broadcast_parameters()
    loader = SyntheticStereo(args.synthetic, args.batch_size, args.crop_height, args.crop_width, args.max_disp, seed=rank)
    for epoch in range(args.start_epoch, args.epochs):
        g_optimizer.set_lr(args.lr * (0.5 ** sum(epoch >= m for m in args.milestones)))
        train_loss = train(loader, m_model, fix_model, g_optimizer, epoch, args, device)
        if rank == 0:
            save_checkpoint({"epoch": epoch + 1, "m_model": args.m_model, "state_dict": m_model.state_dict(),
                             "best_rmse": -1}, False, args.save_path)
            print(f"epoch {epoch}: train loss {train_loss:.5f}")


if __name__ == "__main__":
    main()
