#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv3x3_col -c 2 -o /tmp/r2j_col -f python tools/conv_layers.py --layers "deconv1" --ops fwd,dgrad --iters 1 > gpurun_out/r2j_col.log 2>&1; echo "rc=$?"
ncu -i /tmp/r2j_col.ncu-rep --page raw --csv > gpurun_out/r2j_col.raw.csv 2>/dev/null
ncu -i /tmp/r2j_col.ncu-rep --page source --csv --kernel-id :::1 > gpurun_out/r2j_col.source.csv 2>/dev/null
ls -la gpurun_out/r2j_col*
