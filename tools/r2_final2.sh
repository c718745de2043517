#!/bin/bash
# round-2 closing evidence session: tests, ncu of the new kernels, launch lists, full bench N = 1, reference arm, smoke
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
echo "== full gpu suite"; timeout 1500 python -m pytest tests -m gpu -q --no-header > gpurun_out/r2q_gputests.log 2>&1; echo "rc=$?"; tail -3 gpurun_out/r2q_gputests.log | cut -c1-200
echo "== ncu new kernels"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"stem_mma|stem_wgrad|wgrad_up2|conv3x3_wgrad_kernel|wgrad_multi|pack_dgrad_flat|upsample_nearest_rows|maxpool2_bwd_win" -c 14 -o /tmp/r2q_new -f python tools/r2_kernels.py > gpurun_out/r2q_new.log 2>&1; echo "rc=$?"
ncu -i /tmp/r2q_new.ncu-rep --page raw --csv > gpurun_out/r2q_new.raw.csv 2>/dev/null
echo "== launch lists"
for wl in stage1 stage2 test; do
  timeout 900 ncu --metrics gpu__time_duration.sum,launch__grid_size,launch__block_size --clock-control none -c 6000 --csv --log-file gpurun_out/r2q_launches_$wl.csv python tools/profile_step.py $wl 2 > /dev/null 2>&1; echo "rc=$?"
done
echo "== full bench"; (time timeout 1500 python bench.py > gpurun_out/r2q_bench_n1.json 2> gpurun_out/r2q_bench_n1.err); echo "rc=$?"; cut -c1-400 gpurun_out/r2q_bench_n1.json
echo "== reference arm"; timeout 600 python bench.py --impl reference --steps 20 --warmup 5 2>/dev/null | tail -1 > gpurun_out/r2q_bench_reference.json; cut -c1-300 gpurun_out/r2q_bench_reference.json
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1 | cut -c1-300
du -sh gpurun_out
