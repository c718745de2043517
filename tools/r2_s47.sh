#!/bin/bash
for c in 96 1000 4000 100000 96 1000 4000 100000; do
  echo -n "FALN_WGRAD_BATCH_CHUNKS=$c  "
  FALN_WGRAD_BATCH_CHUNKS=$c timeout 600 python bench.py --steps 100 --no-extras --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
r=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('stage1', round(r['ms_per_step'],4))
"
done
