#!/bin/bash
timeout 600 python -m pytest tests/test_wgrad_gpu.py tests/test_model_gpu.py tests/test_reference_gpu.py -m gpu -x -q 2>&1 | tail -3
for v in 1000 0 1000 0; do
  # FALN_WGRAD_UP2_BATCH=0 keeps the folded deconv weight gradients as separate launches (A/B switch read by backbone.py)
  echo -n "FALN_WGRAD_UP2_BATCH=$v  "
  FALN_WGRAD_UP2_BATCH=$v timeout 600 python bench.py --steps 100 --no-extras --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
r=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('stage1', round(r['ms_per_step'],4))
"
done
