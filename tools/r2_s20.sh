#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_conv_gpu.py tests/test_metrics.py tests/test_model_gpu.py tests/test_reference_gpu.py -m gpu -x -q 2>&1 | tail -8
for v in 0 1 0 1; do
  echo "== FALN_CONV_SPLITK=$v"
  FALN_CONV_SPLITK=$v timeout 600 python bench.py --steps 100 --extras stage2,test --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
r=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('stage1', round(r['ms_per_step'],4), 'stage2', round(r['extras']['stage2']['ms_per_step'],4), 'test', round(r['extras']['test']['ms_per_step'],4), 'conv frac', r['roofline']['frac'])
"
done
