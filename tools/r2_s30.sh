#!/bin/bash
# per-bucket Adam A/B at N = 1, then the lagged-readback test + DDP-related GPU tests
for v in 0 1 0 1; do
  echo "== FALN_BUCKET_ADAM=$v"
  FALN_BUCKET_ADAM=$v timeout 600 python bench.py --steps 100 --no-extras --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
r=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('stage1', round(r['ms_per_step'],4), 'e2e', round(r['e2e']['ms_per_step'],4))
"
done
for v in 0 1; do
  echo "== stage2 FALN_BUCKET_ADAM=$v"
  FALN_BUCKET_ADAM=$v timeout 600 python bench.py --workload stage2 --steps 50 --no-extras --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
r=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('stage2', round(r['ms_per_step'],4), 'e2e', round(r['e2e']['ms_per_step'],4))
"
done
timeout 900 python -m pytest tests/test_model_gpu.py tests/test_kernels_gpu.py -m gpu -x -q 2>&1 | tail -3
