#!/bin/bash
for v in 0 1 0 1; do
  echo -n "FALN_DEBUG_SKIP_BIAS_SUMS=$v  "
  FALN_DEBUG_SKIP_BIAS_SUMS=$v timeout 600 python bench.py --steps 100 --no-extras --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
r=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('stage1', round(r['ms_per_step'],4))
"
done
