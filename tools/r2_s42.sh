#!/bin/bash
for n in 4 6 8; do
echo "== FALN_STEM_CTAS=$n"
FALN_STEM_CTAS=$n timeout 300 python - <<'PY'
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
    print(f"stem {B}x{H}x{W} -> {Cout}: mma {t(lambda: CN.stem_conv(x, w, b, act)):7.1f} us")
PY
done
timeout 900 python -m pytest tests/test_model_gpu.py tests/test_reference_gpu.py tests/test_variants.py tests/test_conv_gpu.py -m gpu -x -q 2>&1 | tail -3
for v in 1 0; do
  for wl in stage1 stage2 test; do
  echo -n "FALN_STEM_FMA=$v $wl "
  FALN_STEM_FMA=$v timeout 600 python bench.py --workload $wl --steps 50 --no-extras --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
r=json.loads(sys.stdin.read().strip().splitlines()[-1])
print(round(r['ms_per_step'],4))
"
  done
done
