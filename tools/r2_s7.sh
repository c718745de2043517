#!/bin/bash
# round-2 session 7: programmatic dependent launch A/B
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
echo "== conv/model tests (PDL on)"; timeout 900 python -m pytest tests/test_conv_gpu.py tests/test_wgrad_gpu.py tests/test_model_gpu.py tests/test_med_gpu.py -m gpu -q --no-header 2>&1 | tail -4
for pdl in 1 0 1 0; do
  echo "== bench stage1 FALN_PDL=$pdl"; FALN_PDL=$pdl timeout 600 python bench.py --steps 100 --no-extras --no-cpu-baseline 2>/dev/null | python -c "import sys,json; r=json.loads(sys.stdin.read()); print('stage1', r['value'], r['ms_per_step'], r['e2e']['value'], r['gpu_launches'])"
done
for pdl in 1 0; do
  echo "== bench stage2 FALN_PDL=$pdl"; FALN_PDL=$pdl timeout 600 python bench.py --workload stage2 --steps 40 --no-extras --no-cpu-baseline 2>/dev/null | python -c "import sys,json; r=json.loads(sys.stdin.read()); print('stage2', r['value'], r['ms_per_step'], r['e2e']['value'])"
  echo "== bench test FALN_PDL=$pdl"; FALN_PDL=$pdl timeout 600 python bench.py --workload test --steps 40 --no-extras --no-cpu-baseline 2>/dev/null | python -c "import sys,json; r=json.loads(sys.stdin.read()); print('test', r['value'], r['ms_per_step'], r['e2e']['value'])"
done
