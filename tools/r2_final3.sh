#!/bin/bash
# last check of round 2: full GPU suite, full bench N = 1, Stage-1 launch list, smoke
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
echo "== full gpu suite"; timeout 900 python -m pytest tests -m gpu -q --no-header > gpurun_out/r2r_gputests.log 2>&1; echo "rc=$?"; tail -2 gpurun_out/r2r_gputests.log | cut -c1-200
echo "== full bench"; timeout 900 python bench.py > gpurun_out/r2r_bench_n1.json 2> gpurun_out/r2r_bench_n1.err; echo "rc=$?"; cut -c1-300 gpurun_out/r2r_bench_n1.json
echo "== launch list"; timeout 300 ncu --metrics gpu__time_duration.sum,launch__grid_size,launch__block_size --clock-control none -c 3000 --csv --log-file gpurun_out/r2r_launches_stage1.csv python tools/profile_step.py stage1 2 > /dev/null 2>&1; echo "rc=$?"
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1 | cut -c1-200
