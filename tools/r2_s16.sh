#!/bin/bash
set -u
cd "$(dirname "$0")/.."
echo "== conv tests"; timeout 900 python -m pytest tests/test_conv_gpu.py tests/test_model_gpu.py -m gpu -q --no-header 2>&1 | tail -3 | cut -c1-200
L="conv2_1.*,iconv3,conv3_1.*,iconv4,conv4_1.*,conv5_1.*,conv6_1.*,deconv6,iconv5"
echo "== layers (graph-timed)"; timeout 300 python tools/conv_layers.py --time --graph --iters 50 --layers "$L" --ops fwd,dgrad 2>&1 | tail -20
run() { timeout 400 python bench.py --steps 60 --no-extras --no-cpu-baseline "$@" 2>/dev/null | python -c "import sys,json; r=json.loads(sys.stdin.read()); print('   ', r['ms_per_step'], r['e2e']['ms_per_step'])"; }
echo stage1; run; run
echo stage2; run --workload stage2
echo test; run --workload test
