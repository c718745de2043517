#!/bin/bash
# split-K cluster conv kernel + cluster percentile: parity, then A/B timing
mkdir -p gpurun_out
export FALN_DEBUG=1
timeout 600 python -m pytest tests/test_conv_gpu.py -x -q 2>&1 | grep -v "^conv3x3_\|^$" | tail -15
unset FALN_DEBUG
timeout 300 python -m pytest tests/test_metrics.py -m gpu -x -q 2>&1 | tail -5
L="conv4.0,conv4_1.*,conv5.0,conv5_1.*,conv6.0,conv6_1.*,deconv6,iconv6,deconv5,iconv5,conv3_1.*,iconv4"
echo "== split-K off"; FALN_CONV_SPLITK=0 timeout 300 python tools/conv_layers.py --time --graph --iters 20 --ops fwd,dgrad --layers "$L" 2>&1 | tail -30
echo "== split-K on"; FALN_DEBUG=1 timeout 300 python tools/conv_layers.py --time --graph --iters 20 --ops fwd,dgrad --layers "$L" 2>&1 | grep -v "conv3x3_tc_kernel\|row_kernel" | tail -50
