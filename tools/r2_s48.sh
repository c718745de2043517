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
run FALN_WGRAD_MIN_CHUNKS=8
run FALN_WGRAD_MIN_CHUNKS=32
run FALN_WGRAD_FILL_PCT=125
run FALN_WGRAD_FILL_PCT=150
run A=0
run FALN_WGRAD_MIN_CHUNKS=8
run FALN_WGRAD_FILL_PCT=125
