#!/usr/bin/env python
"""List the library (non-faln / non-m3) kernels of one eager step with the Python frames that launched them.

    python tools/aten_residue.py stage1|stage2|test
"""
import collections
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch.profiler import ProfilerActivity, profile

from fal_net_b200 import models, steps
from fal_net_b200.trainer import FlatAdamDDP

wl = sys.argv[1] if len(sys.argv) > 1 else "stage1"
dev = torch.device("cuda:0")
torch.manual_seed(0)
model = models.FAL_netB(no_levels=49).to(dev)
B, H, W = (8, 192, 640) if wl != "test" else (8, 375, 1242)
g = torch.Generator().manual_seed(1)
left = (torch.rand(B, 3, H, W, generator=g) - 0.43).to(dev)
right = (torch.rand(B, 3, H, W, generator=g) - 0.43).to(dev)
mx = torch.full((B, 1, 1), 300.0, device=dev)
mn = mx * 2 / 300
opt = FlatAdamDDP(model, lr=1e-4) if wl != "test" else None
fix = None
if wl == "stage2":
    torch.manual_seed(1)
    fix = models.FAL_netB(no_levels=49).to(dev).eval()
    for p_ in fix.parameters():
        p_.requires_grad_(False)          # frozen teacher (Train_Stage2_K.py): its weight packs are cached, not refreshed per step


def step():
    if wl == "test":
        steps.test_disp(model, left, mn, mx, f_post_process=True)
        return
    opt.zero_grad()
    if wl == "stage1":
        loss = steps.stage1_loss(model, left, right, mn, mx, a_p=0.0)[0]
    else:
        loss = steps.stage2_loss(model, fix, left, right, mn, mx, a_p=0.01)["loss"]
    loss.backward()
    opt.step()


for _ in range(3):
    step()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA], with_stack=True) as prof:
    step()
    torch.cuda.synchronize()
# map kernel launches back to the launching op through the correlation id
evs = prof.events()
by_stack = collections.Counter()
n_ours = 0
for e in evs:
    if e.device_type != torch.autograd.DeviceType.CUDA:
        continue
    name = e.name
    if "faln::" in name or "m3::" in name:
        n_ours += 1
        continue
for e in evs:
    if e.device_type == torch.autograd.DeviceType.CPU and e.kernels:
        ks = [k.name for k in e.kernels if not ("faln::" in k.name or "m3::" in k.name)]
        if not ks:
            continue
        frames = [f for f in (e.stack or []) if "fal_net_b200/" in f][:3]
        by_stack[(e.name, " <- ".join("fal_net_b200/" + s.split("fal_net_b200/")[-1] for s in frames))] += len(ks)
print(f"# {wl}: our kernels {n_ours}; library launches by (op, repo frames):")
for (op, st), c in sorted(by_stack.items(), key=lambda kv: -kv[1]):
    print(f"{c:4d}  {op:40s} {st}")
