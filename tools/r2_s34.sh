#!/bin/bash
timeout 600 python -m pytest tests/test_wgrad_gpu.py -m gpu -x -q 2>&1 | tail -4
timeout 300 python - <<'PY'
import torch, sys
sys.path.insert(0, '.')
from fal_net_b200 import conv_native as CN
CL = torch.channels_last
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
for name, cin, cout, H, W in (("deconv1", 64, 64, 96, 320), ("deconv2", 128, 64, 48, 160), ("deconv3", 256, 128, 24, 80),
                              ("deconv4", 256, 128, 12, 40), ("deconv5", 256, 128, 6, 20), ("deconv6", 512, 256, 3, 10)):
    B = 8
    h = torch.randn(B, cin, H, W, device=dev).to(torch.bfloat16).contiguous(memory_format=CL)
    g = torch.randn(B, cout, 2 * H, 2 * W, device=dev).to(torch.bfloat16).contiguous(memory_format=CL)
    dW = torch.zeros(cout, cin, 3, 3, device=dev).contiguous(memory_format=CL)
    old = t(lambda: CN.conv3x3_wgrad(g, CN.upsample_nearest(h, (2 * H, 2 * W)), dW, cout=cout, cx=cin))
    new = t(lambda: CN.conv3x3_wgrad_up2(g, h, dW, cout=cout, cx=cin))
    print(f"{name}: upsample + plain wgrad {old:7.1f} us   folded {new:7.1f} us")
PY
for v in 1 0 1 0; do
  echo -n "FALN_NO_UP2_WGRAD=$v  "
  FALN_NO_UP2_WGRAD=$v timeout 600 python bench.py --steps 100 --no-extras --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
r=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('stage1', round(r['ms_per_step'],4))
"
done
