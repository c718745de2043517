#!/usr/bin/env python
"""Run a few steps of one bench workload (for `ncu --metrics gpu__time_duration.sum`)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from fal_net_b200 import models, steps, loss_functions as LF
from fal_net_b200.trainer import FlatAdamDDP
wl = sys.argv[1] if len(sys.argv) > 1 else "stage1"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 2
dev = torch.device("cuda:0")
torch.manual_seed(0)
model = models.FAL_netB(no_levels=49).to(dev)
B, H, W = (8, 192, 640) if wl != "test" else (8, 375, 1242)
g = torch.Generator().manual_seed(1)
left = (torch.rand(B, 3, H, W, generator=g) - 0.43).to(dev)
right = (torch.rand(B, 3, H, W, generator=g) - 0.43).to(dev)
mx = torch.full((B, 1, 1), 300.0, device=dev); mn = mx * 2 / 300
if wl == "test":
    for _ in range(n):
        steps.test_disp(model, left, mn, mx, f_post_process=True)
else:
    opt = FlatAdamDDP(model, lr=1e-4)
    fix = None
    if wl == "stage2":
        torch.manual_seed(1); fix = models.FAL_netB(no_levels=49).to(dev).eval()
        for p_ in fix.parameters():
            p_.requires_grad_(False)      # frozen teacher (Train_Stage2_K.py): its weight packs are cached, not refreshed per step
    for _ in range(n):
        opt.zero_grad()
        if wl == "stage1":
            loss = steps.stage1_loss(model, left, right, mn, mx, a_p=0.0)[0]
        else:
            loss = steps.stage2_loss(model, fix, left, right, mn, mx, a_p=0.01)["loss"]
        loss.backward(); opt.step()
torch.cuda.synchronize()
