#!/bin/bash
# round-2 final evidence session, part 2 (after the epilogue / stream changes): conv ncu, launch lists, full bench N=1
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
echo "== full gpu suite"; timeout 1500 python -m pytest tests -m gpu -q --no-header > gpurun_out/r2g_gputests.log 2>&1; echo "rc=$?"; tail -3 gpurun_out/r2g_gputests.log | cut -c1-200
echo "== ncu conv"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:conv3x3 -c 22 -o /tmp/r2g_conv -f python tools/conv_layers.py --layers "deconv1,conv0_1.*,iconv3,conv5_1.*,conv3_1.*,conv2_1.*" --ops fwd,dgrad,wgrad --iters 1 > gpurun_out/r2g_conv.log 2>&1; echo "rc=$?"
ncu -i /tmp/r2g_conv.ncu-rep --page raw --csv > gpurun_out/r2g_conv.raw.csv 2>/dev/null
echo "== launch lists"
timeout 900 ncu --metrics gpu__time_duration.sum,launch__grid_size,launch__block_size --clock-control none -c 3000 --csv --log-file gpurun_out/r2g_launches_stage1.csv python tools/profile_step.py stage1 2 > /dev/null 2>&1; echo "rc=$?"
timeout 900 ncu --metrics gpu__time_duration.sum,launch__grid_size,launch__block_size --clock-control none -c 6000 --csv --log-file gpurun_out/r2g_launches_stage2.csv python tools/profile_step.py stage2 2 > /dev/null 2>&1; echo "rc=$?"
timeout 900 ncu --metrics gpu__time_duration.sum,launch__grid_size,launch__block_size --clock-control none -c 3000 --csv --log-file gpurun_out/r2g_launches_test.csv python tools/profile_step.py test 2 > /dev/null 2>&1; echo "rc=$?"
echo "== full bench"; (time timeout 1200 python bench.py > gpurun_out/r2g_bench_n1.json 2> gpurun_out/r2g_bench_n1.err); echo "rc=$?"; cut -c1-300 gpurun_out/r2g_bench_n1.json
echo "== reference arm"; timeout 600 python bench.py --impl reference --steps 20 --warmup 5 2>/dev/null | tail -1 > gpurun_out/r2g_bench_reference.json; cut -c1-200 gpurun_out/r2g_bench_reference.json
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1 | cut -c1-300
du -sh gpurun_out
