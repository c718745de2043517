#!/bin/bash
timeout 900 python -m pytest tests/test_wgrad_gpu.py -m gpu -x -q 2>&1 | tail -4
L="conv0_1.*"
for v in 1 0; do
  if [ $v = 1 ]; then export FALN_WGRAD_NO_HALO32=1; else unset FALN_WGRAD_NO_HALO32; fi
  echo "== NO_HALO32=$v"
  timeout 300 python tools/conv_layers.py --time --graph --iters 20 --ops wgrad --layers "$L" 2>&1 | tail -1
  for i in 1 2; do
  timeout 600 python bench.py --steps 100 --no-extras --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
r=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('stage1', round(r['ms_per_step'],4))
"
  done
done
