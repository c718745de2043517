#!/usr/bin/env python
"""Launch the kernels added in the second half of round 2 once each at Stage-1 / Stage-2 shapes (for `ncu --set full`):
fused tensor-core stem, folded deconv weight gradient, weight gradient with the bias gradient in its spare slot, flat
data-gradient re-pack, row-structured up-sampling, window-based max-pool backward."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from fal_net_b200 import _lib, conv_native as CN, models
from fal_net_b200.trainer import FlatAdamDDP

dev = torch.device("cuda:0")
CL = torch.channels_last
g = torch.Generator(device=dev).manual_seed(3)


def rnd(*shape):
    return torch.randn(*shape, device=dev, generator=g).to(torch.bfloat16).contiguous(memory_format=CL)


B = 8
img = torch.randn(B, 3, 192, 640, device=dev, generator=g)
for cout, act, bb in ((32, 1, 8), (64, 2, 16)):
    w = torch.randn(cout, 3, 3, 3, device=dev, generator=g) * 0.3
    CN.stem_conv(img if bb == 8 else torch.cat((img, img)), w, torch.zeros(cout, device=dev), act)
for cin, cout, H, W in ((64, 64, 96, 320), (256, 128, 24, 80)):          # deconv1, deconv3 (low-resolution input sizes)
    h, gy = rnd(B, cin, H, W), rnd(B, cout, 2 * H, 2 * W)
    dW = torch.zeros(cout, cin, 3, 3, device=dev).contiguous(memory_format=CL)
    CN.conv3x3_wgrad_up2(gy, h, dW, cout=cout, cx=cin)
for cin, cout, H, W, stride in ((64, 64, 96, 320, 1), (64, 128, 96, 320, 2), (32, 32, 192, 640, 1)):
    x = rnd(B, cin, H, W)
    gy = rnd(B, cout, (H - 1) // stride + 1, (W - 1) // stride + 1)
    dW = torch.zeros(cout, cin, 3, 3, device=dev).contiguous(memory_format=CL)
    CN.conv3x3_wgrad(gy, x, dW, cout=cout, cx=cin, stride=stride, dbias=torch.zeros(cout, device=dev))
# the stem's weight + bias gradient from the image
dW0 = torch.zeros(32, 3, 3, 3, device=dev).contiguous(memory_format=CL)
CN.stem_wgrad(img, rnd(B, 32, 192, 640), dW0, torch.zeros(32, device=dev))
torch.manual_seed(0)
opt = FlatAdamDDP(models.FAL_netB(no_levels=49).to(dev), lr=1e-4)
opt._repack_dgrad()
CN.upsample_nearest(rnd(B, 64, 188, 621), (375, 1242))
CN.maxpool2_bwd(rnd(16, 64, 192, 640), rnd(16, 64, 96, 320), dact=2)
torch.cuda.synchronize()
print("ok")
