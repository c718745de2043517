#!/bin/bash
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum,launch__grid_size,launch__block_size --clock-control none -c 3000 --csv --log-file gpurun_out/r2k_launches_stage1.csv python tools/profile_step.py stage1 2 > /dev/null 2>&1; echo "rc=$?"
