#!/bin/bash
mkdir -p gpurun_out
FALN_DEBUG=1 timeout 600 python -m pytest tests/test_conv_gpu.py -x -q -k "bf16_nhwc or dgrad or planar" 2>&1 | grep -v "^conv3x3_tc\|^conv3x3_row\|^conv3x3_splitk\|^$" | tail -25
L="conv0_1.*,conv1_1.*,iconv2,iconv1+conv0 (folded),deconv1,deconv2"
echo "== col off"; FALN_CONV_COL=0 timeout 300 python tools/conv_layers.py --time --graph --iters 20 --ops fwd,dgrad --layers "$L" 2>&1 | tail -12
echo "== col on"; timeout 300 python tools/conv_layers.py --time --graph --iters 20 --ops fwd,dgrad --layers "$L" 2>&1 | tail -12
