#!/bin/bash
timeout 900 python -m pytest tests/test_conv_gpu.py tests/test_model_gpu.py -m gpu -x -q 2>&1 | tail -3
for v in flat rows; do
  if [ $v = flat ]; then export FALN_EW_FLAT=1; else unset FALN_EW_FLAT; fi
  echo -n "$v  "
  timeout 600 python bench.py --workload test --steps 30 --no-extras --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
r=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('test', round(r['ms_per_step'],4))
"
  echo -n "$v  "
  timeout 600 python bench.py --workload stage2 --steps 50 --no-extras --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
r=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('stage2', round(r['ms_per_step'],4))
"
done
