#!/bin/bash
timeout 900 python -m pytest tests/test_model_gpu.py tests/test_reference_gpu.py tests/test_variants.py -m gpu -x -q 2>&1 | tail -4
for v in 1 0 1 0; do
  echo -n "FALN_NO_FUSED_BIAS_GRAD=$v  "
  FALN_NO_FUSED_BIAS_GRAD=$v timeout 600 python bench.py --steps 100 --no-extras --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
r=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('stage1', round(r['ms_per_step'],4))
"
done
echo -n "stage2 fused  "; timeout 600 python bench.py --workload stage2 --steps 50 --no-extras --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
r=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('stage2', round(r['ms_per_step'],4))
"
