#!/usr/bin/env python
"""Inference entry point on the B200-native hot path.

Mirrors /root/reference/Test_KITTI.py: flag names (:34-60), ``validate()`` with flip (-fpp, :199-203) or multi-scale
(-mspp, the reference's default, :204-205,287-300) post-processing.  Images are sharded across ranks (one process per
GPU, no collective).  Metrics / PNG / PLY dumps (:211-280) are host-side bookkeeping and out of scope."""
import argparse
import time

import torch

from fal_net_b200 import models, steps
from fal_net_b200.entry_common import SyntheticStereo, init_distributed

parser = argparse.ArgumentParser(description="Testing pan generation (B200-native hot path)",
                                 formatter_class=argparse.ArgumentDefaultsHelpFormatter)
def _flag(v):
    """The reference declares its switches as plain valued options (``-fpp True``); accept that form and the bare switch."""
    return str(v).lower() not in ("0", "false", "no", "none", "")


parser.add_argument("-d", "--data", default="")
parser.add_argument("-tn", "--tdataName", default="Kitti_eigen_test_improved")
parser.add_argument("-relbase", "--rel_baselne", type=float, default=1)
parser.add_argument("-mdisp", "--max_disp", type=float, default=300)
parser.add_argument("-mindisp", "--min_disp", type=float, default=2)
parser.add_argument("-b", "--batch_size", type=int, default=8, help="reference default: 1 (its loader is b1, :113)")
parser.add_argument("-m", "--model", default="FAL_netB")
parser.add_argument("-no_levels", "--no_levels", type=int, default=49)
parser.add_argument("--checkpoint", default=None, help="checkpoint.pth.tar of the reference or of this repo")
parser.add_argument("-fpp", "--f_post_process", nargs="?", const=True, default=False, type=_flag)
parser.add_argument("-mspp", "--ms_post_process", nargs="?", const=True, default=True, type=_flag,
                    help="multi-scale post-processing (the reference's shipped default, :57-58); -fpp takes precedence")
# accepted for command-line compatibility with the reference (:34-60); data loading, metric tables and PNG / PLY dumps are
# out of scope, the values are not used
for _opts, _dflt in ((("-eval", "--evaluate"), True), (("-save", "--save"), False), (("-save_pc", "--save_pc"), False),
                     (("-save_pan", "--save_pan"), False), (("-save_input", "--save_input"), False),
                     (("-w", "--workers"), 4), (("--print-freq", "-p"), 10), (("-gpu_no", "--gpu_no"), "0"),
                     (("-dt", "--dataset"), "Kitti_stage2"), (("-ts", "--time_stamp"), ""), (("-dtl", "--details"), ""),
                     (("-median", "--median"), False)):
    parser.add_argument(*_opts, default=_dflt, help="accepted, unused")
parser.add_argument("--sparse", action="store_true", default=False, help="accepted, unused")
parser.add_argument("--images", type=int, default=64, help="number of synthetic 375x1242 images (Eigen-split shaped)")


@torch.no_grad()
def validate(val_loader, pan_model, args, device):
    right_shift = args.max_disp * args.rel_baselne
    n, t0 = 0, time.time()
    out = []
    for (input_left, _), _ in val_loader:
        input_left = input_left.to(device, non_blocking=True)
        B = input_left.shape[0]
        max_disp = torch.full((B, 1, 1), right_shift, device=device)
        min_disp = max_disp * args.min_disp / args.max_disp
        disp = steps.test_disp(pan_model, input_left, min_disp, max_disp, f_post_process=args.f_post_process,
                               ms_post_process=args.ms_post_process and not args.f_post_process)
        out.append(disp)
        n += B
    torch.cuda.synchronize()
    return out, n, time.time() - t0


def main(argv=None):
    args = parser.parse_args(argv)
    rank, world, device = init_distributed()
    data = torch.load(args.checkpoint, map_location="cpu") if args.checkpoint else None
    pan_model = models.__dict__[args.model](data, no_levels=args.no_levels).to(device).eval()
    per_rank = (args.images + world - 1) // world
    loader = SyntheticStereo((per_rank + args.batch_size - 1) // args.batch_size, args.batch_size, 375, 1242, args.max_disp,
                             seed=1000 + rank)
    _, n, dt = validate(loader, pan_model, args, device)
    print(f"rank {rank}: {n} images in {dt:.3f} s ({n / dt:.1f} images/s)")


if __name__ == "__main__":
    main()
