#!/bin/bash
for wl in stage2 stage1 test; do timeout 300 python tools/aten_residue.py $wl 2>&1 | grep -v Warn | tail -60; done
