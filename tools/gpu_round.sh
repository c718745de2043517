#!/bin/bash
# One GPU-box visit: parity tests, bench lines, MED microbench, ncu launch list, ncu full capture of the MED kernels.
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_stage1.json 2> gpurun_out/bench_stage1.err
tail -c 3000 gpurun_out/bench_stage1.json
timeout 300 python tools/bench_med.py > gpurun_out/bench_med.jsonl 2> gpurun_out/bench_med.err
cat gpurun_out/bench_med.jsonl
timeout 600 ncu --metrics gpu__time_duration.sum,launch__grid_size,launch__block_size --clock-control none -c 3000 --csv --log-file gpurun_out/launches_stage1.csv python tools/profile_step.py stage1 2 > gpurun_out/ncu_launch.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:med_ -c 8 -f -o gpurun_out/med_full python tools/bench_med.py --profile 8,49,375,1242 > gpurun_out/ncu_med.log 2>&1
tail -3 gpurun_out/ncu_med.log
