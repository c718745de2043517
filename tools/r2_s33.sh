#!/bin/bash
for mc in 4 16 32 64 4 16 32 64 12 24; do
  echo -n "FALN_WGRAD_MIN_CHUNKS=$mc  "
  FALN_WGRAD_MIN_CHUNKS=$mc timeout 600 python bench.py --steps 100 --no-extras --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
r=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('stage1', round(r['ms_per_step'],4))
"
done
for mc in 4 16 32; do
  echo -n "stage2 FALN_WGRAD_MIN_CHUNKS=$mc  "
  FALN_WGRAD_MIN_CHUNKS=$mc timeout 600 python bench.py --workload stage2 --steps 50 --no-extras --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
r=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('stage2', round(r['ms_per_step'],4))
"
done
