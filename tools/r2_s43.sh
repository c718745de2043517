#!/bin/bash
run() {
  echo -n "$* : "
  env "$@" timeout 600 python bench.py --steps 100 --no-extras --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
r=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('stage1', round(r['ms_per_step'],4))
"
}
run A=0
run FALN_WGRAD_UP2_MIN_CHUNKS=2
run FALN_WGRAD_UP2_MIN_CHUNKS=8
run FALN_WGRAD_UP2_MIN_CHUNKS=16
run FALN_WGRAD_FILL_PCT=75
run FALN_WGRAD_FILL_PCT=125
run FALN_WGRAD_FILL_PCT=50
run A=0
run FALN_WGRAD_SMEM_KB=100
run FALN_WGRAD_SMEM_KB=150
run FALN_MAIN_PRIORITY=1
