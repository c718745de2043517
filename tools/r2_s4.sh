#!/bin/bash
# round-2 session 4: input-pipeline GPU tests; ncu --set full of the conv kernels on representative layers; stage-1 launch list
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
echo "== pipeline tests"; timeout 600 python -m pytest tests/test_input_pipeline.py -m gpu -q --no-header 2>&1 | tail -5
echo "== conv layer timing"; timeout 300 python tools/conv_layers.py --time --iters 10 2>&1 | tail -80
echo "== ncu conv full"; timeout 900 ncu --set full --clock-control none --import-source on -k regex:conv3x3 -c 12 -o gpurun_out/r2s4_conv python tools/conv_layers.py --layers "deconv1,conv0_1.*,iconv3,conv5_1.*" --ops fwd,dgrad,wgrad --iters 1 > gpurun_out/r2s4_ncu.log 2>&1; echo "rc=$?"; tail -3 gpurun_out/r2s4_ncu.log
echo "== launch list"; timeout 900 ncu --metrics gpu__time_duration.sum,launch__grid_size,launch__block_size --clock-control none -c 3000 --csv --log-file gpurun_out/r2s4_launches_stage1.csv python tools/profile_step.py stage1 2 > gpurun_out/r2s4_launches.log 2>&1; echo "rc=$?"
ls -la gpurun_out/r2s4*
