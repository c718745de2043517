#!/bin/bash
# round-2 session 9: up-sample folded deconv: parity, A/B timing
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
echo "== up2 unit tests"; timeout 600 python -m pytest tests/test_conv_gpu.py -m gpu -q --no-header -k upsample_folded 2>&1 | tail -8 | cut -c1-300
echo "== model + reference tests"; timeout 1200 python -m pytest tests/test_model_gpu.py tests/test_reference_gpu.py tests/test_variants.py -m gpu -q --no-header 2>&1 | tail -6 | cut -c1-300
for v in 0 1 0 1; do
  echo "== bench stage1 FALN_NO_UP2=$v"; FALN_NO_UP2=$v timeout 600 python bench.py --steps 100 --no-extras --no-cpu-baseline 2>/dev/null | python -c "import sys,json; r=json.loads(sys.stdin.read()); print('stage1', r['value'], r['ms_per_step'], r['e2e']['value'], r['gpu_launches'], r['roofline']['frac'], r['roofline']['frac_of_layerwise_roofline_ss'])"
done
for v in 0 1; do
  echo "== bench stage2 FALN_NO_UP2=$v"; FALN_NO_UP2=$v timeout 600 python bench.py --workload stage2 --steps 40 --no-extras --no-cpu-baseline 2>/dev/null | python -c "import sys,json; r=json.loads(sys.stdin.read()); print('stage2', r['value'], r['ms_per_step'], r['e2e']['value'])"
done
