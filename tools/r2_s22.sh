#!/bin/bash
for h in none small all none small; do
  echo "== FALN_HACK_SKIP_WGRAD=$h"
  FALN_HACK_SKIP_WGRAD=$h timeout 600 python bench.py --steps 100 --no-extras --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
r=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('stage1', round(r['ms_per_step'],4))
"
done
python - <<'P'
import torch
from fal_net_b200 import postproc
d = 120 * torch.rand(8, 1, 375, 1242, device="cuda")
for _ in range(3):
    postproc.percentile_rows(d, 95.0, add=1e-6)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(20):
    postproc.percentile_rows(d, 95.0, add=1e-6)
e1.record(); torch.cuda.synchronize()
print("percentile us", e0.elapsed_time(e1) / 20 * 1e3)
P
timeout 300 python -m pytest tests/test_metrics.py -m gpu -x -q 2>&1 | tail -2
