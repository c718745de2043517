#!/bin/bash
timeout 300 python -m pytest tests/test_conv_gpu.py -x -q 2>&1 | tail -2
L="conv0_1.*,conv1_1.*,deconv1,iconv1+conv0 (folded)"
timeout 300 python tools/conv_layers.py --time --graph --iters 20 --ops fwd,dgrad --layers "$L" 2>&1 | tail -8
FALN_COL_DBG=1 timeout 120 python tools/conv_layers.py --time --iters 3 --ops fwd --layers "deconv1" 2>&1 | tail -26
