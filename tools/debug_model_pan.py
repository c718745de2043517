#!/usr/bin/env python
"""Isolate MED-kernel error from bf16 conv noise for the whole-model forward (tests/test_model_gpu.py)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from fal_net_b200 import models, med
from oracle import falnet_oracle as O
from tests.helpers import disp_range, images, rel_err, rel_l2
dev = torch.device("cuda:0")
for H, W in [(48, 160), (50, 166), (64, 192)]:
    for seed in (0, 1, 2):
        torch.manual_seed(seed)
        m = models.FAL_netB(None, no_levels=49).to(dev)
        p = {k: v.detach().cpu().clone() for k, v in m.state_dict().items()}
        B = 2
        left = images(B, H, W, 1234)
        mn, mx = disp_range(B)
        with torch.no_grad():
            lg = m.logits(left.to(dev), mx.to(dev))
            pan, disp, mL, mR = med.med_section(lg, left.to(dev), mn.to(dev), mx.to(dev), True, True, True)
        rp, rd, rmL, rmR = O.falnet_forward(p, left, mn, mx, True, True, True)
        lgc = lg[..., :W].cpu().contiguous()
        d, xo = O.level_tables(mn, mx, 49, W)
        ref = O.med_forward_closed(lgc, left, d, xo)
        flow = torch.ones(B, 1, H, W) * (mx.view(B, 1, 1, 1) / 100)
        rl = torch.nn.functional.conv2d(O.backbone_forward(p, left, flow), p["conv0.weight"], p["conv0.bias"])
        print(H, W, seed, "pan vs full oracle: max %.3e l2 %.3e | pan vs oracle-MED-on-our-logits: %.3e | disp %.3e / %.3e"
              % (rel_err(pan, rp), rel_l2(pan, rp), rel_err(pan, ref["pan"]), rel_err(disp, rd), rel_err(disp, ref["disp"])),
              "" if rl is None else "logits max %.3e l2 %.3e absmax %.2f" % (rel_err(lgc, rl), rel_l2(lgc, rl), float(rl.abs().max())))
