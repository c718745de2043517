#!/bin/bash
# One gpurun session: MED parity, MED microbench (third vs second generation, tuning sweeps), ncu capture, full GPU suite,
# bench.py.  Everything is logged under gpurun_out/; each step has its own timeout so a hang cannot eat the box.
set -u
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
TAG=${1:-s}
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/${TAG}_smi.txt 2>&1
echo "== med tests"; timeout 400 python -m pytest tests/test_med_gpu.py -q > gpurun_out/${TAG}_medtests.log 2>&1; echo "rc=$?"; tail -4 gpurun_out/${TAG}_medtests.log
echo "== bench_med v3";  timeout 200 python tools/bench_med.py --quick > gpurun_out/${TAG}_med_v3.jsonl 2> gpurun_out/${TAG}_med_v3.err; echo "rc=$?"; cat gpurun_out/${TAG}_med_v3.jsonl
for cfg in "CTAS=1" "NBUF=1"; do
  envs=""; for kv in $cfg; do envs="$envs FALN_MED3_$kv"; done
  f="gpurun_out/${TAG}_med_sweep_$(echo $cfg | tr ' =' '__').jsonl"
  echo "== sweep $cfg"; env $envs timeout 120 python tools/bench_med.py --quick > "$f" 2>&1; cat "$f" | cut -c1-420
done
echo "== ncu 640"
timeout 300 ncu --set full --import-source on --clock-control none -k regex:med3_ -c 6 -f -o gpurun_out/${TAG}_med3_640 \
  python tools/bench_med.py --profile 16,49,192,640 > gpurun_out/${TAG}_ncu640.log 2>&1; echo "rc=$?"
echo "== ncu 1242"
timeout 300 ncu --set full --import-source on --clock-control none -k regex:med3_ -c 6 -f -o gpurun_out/${TAG}_med3_1242 \
  python tools/bench_med.py --profile 8,49,375,1242 > gpurun_out/${TAG}_ncu1242.log 2>&1; echo "rc=$?"
echo "== full gpu suite"; timeout 900 python -m pytest tests -m gpu -q > gpurun_out/${TAG}_gputests.log 2>&1; echo "rc=$?"; tail -5 gpurun_out/${TAG}_gputests.log
echo "== bench"; timeout 400 python bench.py > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "rc=$?"; cut -c1-600 gpurun_out/${TAG}_bench.json
echo "== bench_med full"; timeout 300 python tools/bench_med.py > gpurun_out/${TAG}_med_full.jsonl 2>&1; echo "rc=$?"
ls -la gpurun_out | head -40
