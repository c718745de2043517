#!/bin/bash
# One gpurun session: MED parity, MED microbench (third vs second generation, tuning sweeps), full GPU suite, bench.py.
# Everything is logged under gpurun_out/; each step has its own timeout so a hang cannot eat the box.
set -u
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/s_smi.txt 2>&1
echo "== med tests"; timeout 400 python -m pytest tests/test_med_gpu.py -x -q > gpurun_out/s_medtests.log 2>&1; echo "rc=$?"; tail -3 gpurun_out/s_medtests.log
echo "== bench_med v3";  timeout 200 python tools/bench_med.py --quick > gpurun_out/s_med_v3.jsonl 2> gpurun_out/s_med_v3.err; echo "rc=$?"; cat gpurun_out/s_med_v3.jsonl
echo "== bench_med v2";  timeout 200 python tools/bench_med.py --quick --flags 32 > gpurun_out/s_med_v2.jsonl 2> gpurun_out/s_med_v2.err; echo "rc=$?"; cat gpurun_out/s_med_v2.jsonl
for cfg in "CTAS=3" "CTAS=2" "G=2 S=4" "G=1 S=5" "G=4 S=2"; do
  envs=""; for kv in $cfg; do envs="$envs FALN_MED3_$kv"; done
  echo "== sweep $cfg"; env $envs timeout 120 python tools/bench_med.py --quick > "gpurun_out/s_med_sweep_$(echo $cfg | tr ' =' '__').jsonl" 2>&1; cat "gpurun_out/s_med_sweep_$(echo $cfg | tr ' =' '__').jsonl"
done
echo "== full gpu suite"; timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/s_gputests.log 2>&1; echo "rc=$?"; tail -5 gpurun_out/s_gputests.log
echo "== bench"; timeout 400 python bench.py > gpurun_out/s_bench.json 2> gpurun_out/s_bench.err; echo "rc=$?"; cat gpurun_out/s_bench.json
