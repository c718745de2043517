#!/bin/bash
echo "== split off"; FALN_CONV_SPLITK=0 timeout 300 python tools/dbg/vgg_split.py 2>&1 | grep -v Warn | tail -20
echo "== split on"; FALN_DEBUG=1 timeout 300 python tools/dbg/vgg_split.py 2>&1 | grep -v "Warn\|tc_kernel\|row_kernel" | tail -30
timeout 300 python -m pytest tests/test_metrics.py -m gpu -x -q 2>&1 | grep -n "assert\|^E" | cut -c1-200 | head
