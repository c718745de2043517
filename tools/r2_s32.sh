#!/bin/bash
L="conv4_1.*,conv5.0,conv5_1.*,conv6.0,conv6_1.*,deconv6,iconv6,deconv5,iconv5,conv4.0,conv3_1.*"
for mc in 4 8 16; do
  echo "== FALN_WGRAD_MIN_CHUNKS=$mc"
  FALN_WGRAD_MIN_CHUNKS=$mc timeout 300 python tools/conv_layers.py --time --graph --iters 20 --ops wgrad --layers "$L" 2>&1 | tail -11
done
for mc in 4 8 16 4 8; do
  echo "== bench FALN_WGRAD_MIN_CHUNKS=$mc"
  FALN_WGRAD_MIN_CHUNKS=$mc timeout 600 python bench.py --steps 100 --no-extras --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
r=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('stage1', round(r['ms_per_step'],4))
"
done
