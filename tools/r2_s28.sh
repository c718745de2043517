#!/bin/bash
# column-walk kernel with sixteen epilogue warps: parity, per-layer A/B, Stage-1 A/B
timeout 300 python -m pytest tests/test_conv_gpu.py -m gpu -x -q -k "column_walk" 2>&1 | tail -2
L="conv0_1.*,deconv2,deconv1,iconv1+conv0 (folded)"
for cfg in "0 8" "1 8" "1 16"; do
  set -- $cfg
  echo "== FALN_CONV_COL=$1 FALN_COL_EW=$2"
  FALN_CONV_COL=$1 FALN_COL_EW=$2 timeout 300 python tools/conv_layers.py --time --graph --iters 20 --ops fwd,dgrad --layers "$L" 2>&1 | tail -8
done
for cfg in "0 8" "1 16"; do
  set -- $cfg
  echo "== bench FALN_CONV_COL=$1 FALN_COL_EW=$2"
  FALN_CONV_COL=$1 FALN_COL_EW=$2 timeout 600 python bench.py --steps 100 --no-extras --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
r=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('stage1', round(r['ms_per_step'],4), 'conv frac', r['roofline']['frac'])
"
done
