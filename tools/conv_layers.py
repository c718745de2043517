#!/usr/bin/env python
"""Launch our conv kernels on chosen FAL_netB layer shapes (for ncu captures and quick A/B timing).

    python tools/conv_layers.py --layers deconv1,conv5_1.*,iconv3 --ops fwd,dgrad,wgrad --iters 3 [--time]
"""
import argparse
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from fal_net_b200 import conv_native as CN  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--layers", default="all")
ap.add_argument("--ops", default="fwd,dgrad,wgrad")
ap.add_argument("--iters", type=int, default=3)
ap.add_argument("--batch", type=int, default=8)
ap.add_argument("--time", action="store_true")
ap.add_argument("--graph", action="store_true", help="time a CUDA-graph replay of the launches (no host launch overhead)")
a = ap.parse_args()
dev = torch.device("cuda:0")
CL = torch.channels_last
g = torch.Generator(device=dev).manual_seed(11)
want = None if a.layers == "all" else set(a.layers.split(","))
B = a.batch
for name, cin, cout, H, W, stride in bench._LAYERS:
    if want is not None and name not in want:
        continue
    Ho, Wo = (H - 1) // stride + 1, (W - 1) // stride + 1
    x = torch.randn(B, cin, H, W, device=dev, generator=g).to(torch.bfloat16).contiguous(memory_format=CL)
    w = torch.randn(cout, cin, 3, 3, device=dev, generator=g) / (3 * cin ** 0.5)
    gy = torch.randn(B, cout, Ho, Wo, device=dev, generator=g).to(torch.bfloat16).contiguous(memory_format=CL)
    wk, wd = CN.pack_weight(w), CN.pack_weight_dgrad(w)
    dW = torch.zeros(cout, 3, 3, cin, device=dev).permute(0, 3, 1, 2)
    parts = [(0, cin)] if (cin == 32 or cin % 64 == 0) else [(0, cin // 64 * 64), (cin // 64 * 64, cin % 64)]
    xs = [x[:, o:o + c].contiguous(memory_format=CL) for o, c in parts] if len(parts) > 1 else [x]
    bias = torch.zeros(cout, device=dev)

    def fwd():
        CN.conv3x3_fwd(x, wk, bias, stride, 1)

    def dgrad():
        for o, c in parts:
            CN.conv3x3_dgrad(gy, wd, (H, W), stride=stride, rows=(o, c))

    def wgrad():
        for (o, c), xp in zip(parts, xs):
            CN.conv3x3_wgrad(gy, xp, dW, cout=cout, cx=c, ci_off=o, stride=stride)
    for op in a.ops.split(","):
        fn = {"fwd": fwd, "dgrad": dgrad, "wgrad": wgrad}[op]
        if a.time and a.graph:
            for _ in range(3):
                fn()
            torch.cuda.synchronize()
            gr = torch.cuda.CUDAGraph()
            with torch.cuda.graph(gr):
                for _ in range(a.iters):
                    fn()
            gr.replay()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            gr.replay()
            e1.record()
            torch.cuda.synchronize()
            print(f"{name:24s} {op:6s} {e0.elapsed_time(e1) / a.iters * 1e3:8.1f} us (graph)")
        elif a.time:
            for _ in range(3):
                fn()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(a.iters):
                fn()
            e1.record()
            torch.cuda.synchronize()
            print(f"{name:24s} {op:6s} {e0.elapsed_time(e1) / a.iters * 1e3:8.1f} us")
        else:
            for _ in range(a.iters):
                fn()
    torch.cuda.synchronize()
