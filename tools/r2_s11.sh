#!/bin/bash
# round-2 session 11: Stage-2 overlap of the frozen pass / VGG labels (A/B), small-layer tile-width experiment
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
echo "== stage2 tests"; timeout 900 python -m pytest tests/test_model_gpu.py tests/test_reference_gpu.py -m gpu -q --no-header -k "stage2 or kslow" 2>&1 | tail -3 | cut -c1-200
for v in 1 0 1 0; do
  echo "== bench stage2 FALN_STAGE2_OVERLAP=$v"; FALN_STAGE2_OVERLAP=$v timeout 600 python bench.py --workload stage2 --steps 40 --no-extras --no-cpu-baseline 2>/dev/null | python -c "import sys,json; r=json.loads(sys.stdin.read()); print('stage2', r['value'], r['ms_per_step'], r['e2e']['value'])"
done
L="conv4_1.*,conv5_1.*,conv6_1.*,deconv6,iconv6,deconv5,iconv5,conv3_1.*"
for pct in 10 50 100 300; do
  echo "== small layers FALN_CONV_NARROW_PCT=$pct"; FALN_CONV_NARROW_PCT=$pct timeout 300 python tools/conv_layers.py --time --iters 20 --layers "$L" --ops fwd,dgrad 2>&1 | tail -16
done
