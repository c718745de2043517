#!/bin/bash
timeout 900 python -m pytest tests/test_model_gpu.py tests/test_variants.py tests/test_conv_gpu.py -m gpu -x -q 2>&1 | tail -3
for v in 1 0 1 0; do
  echo -n "FALN_NO_UP2_PACK_BATCH=$v  "
  FALN_NO_UP2_PACK_BATCH=$v timeout 600 python bench.py --steps 100 --no-extras --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
r=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('stage1', round(r['ms_per_step'],4))
"
done
