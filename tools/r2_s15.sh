#!/bin/bash
set -u
cd "$(dirname "$0")/.."
L="conv2_1.*,iconv3,conv3_1.*,iconv4,iconv2,conv4_1.*,conv5_1.*,conv6_1.*"
echo "== normal"; timeout 300 python tools/conv_layers.py --time --graph --iters 50 --layers "$L" --ops fwd 2>&1 | tail -9
echo "== HACK 77: skip A fill of taps 1..8 (wrong results; timing sensitivity only)"; FALN_HACK=77 timeout 300 python tools/conv_layers.py --time --graph --iters 50 --layers "$L" --ops fwd 2>&1 | tail -9
echo "== HACK 78: skip B (weight) fill except the first (wrong results)"; FALN_HACK=78 timeout 300 python tools/conv_layers.py --time --graph --iters 50 --layers "$L" --ops fwd 2>&1 | tail -9
