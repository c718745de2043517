#!/bin/bash
# round-2 session 10: deferred re-pack parity + bench; per-kernel breakdown
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
echo "== model tests"; timeout 1200 python -m pytest tests/test_model_gpu.py tests/test_variants.py tests/test_kernels_gpu.py -m gpu -q --no-header 2>&1 | tail -4 | cut -c1-300
for i in 1 2; do
echo "== bench stage1"; timeout 600 python bench.py --steps 100 --no-extras --no-cpu-baseline 2>/dev/null > gpurun_out/r2s10_b1.json; python - <<'P'
import json
r=json.loads(open('gpurun_out/r2s10_b1.json').read().strip().splitlines()[-1])
print('stage1', r['value'], r['ms_per_step'], r['e2e']['value'], r['roofline']['frac'], r['roofline']['frac_of_layerwise_roofline_ss'])
for k,v in r['roofline']['by_kernel'].items(): print('   ',k, round(v['ms_per_step'],3), round(v['tflops'],1), v['launches_per_step'])
for k,v in r['roofline_med']['all_med_kernels'].items(): print('   ',k, round(v['avg_ms'],4), round(v['frac'],3))
P
done
echo "== bench stage2"; timeout 600 python bench.py --workload stage2 --steps 40 --no-extras --no-cpu-baseline 2>/dev/null | python -c "import sys,json; r=json.loads(sys.stdin.read()); print('stage2', r['value'], r['ms_per_step'], r['e2e']['value'])"
echo "== bench test"; timeout 600 python bench.py --workload test --steps 40 --no-extras --no-cpu-baseline 2>/dev/null | python -c "import sys,json; r=json.loads(sys.stdin.read()); print('test', r['value'], r['ms_per_step'], r['e2e']['value'])"
