#!/bin/bash
timeout 900 python -m pytest tests/test_conv_gpu.py -m gpu -x -q -k "stem" 2>&1 | tail -15
timeout 300 python - <<'PY'
import torch, sys
sys.path.insert(0, '.')
from fal_net_b200 import conv_native as CN
dev = torch.device('cuda:0')
def t(fn, iters=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(iters): fn()
    g.replay(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); g.replay(); e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e3
for (B, H, W, Cout, act) in ((8, 192, 640, 32, 1), (16, 192, 640, 64, 2), (8, 375, 1242, 32, 1)):
    x = torch.randn(B, 3, H, W, device=dev)
    w = torch.randn(Cout, 3, 3, 3, device=dev) * 0.3
    b = torch.randn(Cout, device=dev)
    CN.STEM_MMA = True
    new = t(lambda: CN.stem_conv(x, w, b, act))
    CN.STEM_MMA = False
    old = t(lambda: CN.stem_conv(x, w, b, act))
    CN.STEM_MMA = True
    mb = (B * 3 * H * W * 4 + B * H * W * Cout * 2) / 1e6
    print(f"stem {B}x{H}x{W} -> {Cout}: fma {old:7.1f} us   mma {new:7.1f} us   ({mb / new / 1e3 * 1e3:.0f} GB/s)")
PY
