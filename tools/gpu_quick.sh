#!/bin/bash
# Short gpurun session: MED parity + microbench (+ optional env sweeps given as arguments "K=V K=V" ...), full suite, bench.
set -u
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
TAG=${1:-q}; shift || true
echo "== med tests"; timeout 400 python -m pytest tests/test_med_gpu.py -q > gpurun_out/${TAG}_medtests.log 2>&1; echo "rc=$?"; tail -3 gpurun_out/${TAG}_medtests.log
echo "== bench_med"; timeout 300 python tools/bench_med.py > gpurun_out/${TAG}_med_full.jsonl 2> gpurun_out/${TAG}_med.err; echo "rc=$?"; cut -c1-330 gpurun_out/${TAG}_med_full.jsonl
for cfg in "$@"; do
  envs=""; for kv in $cfg; do envs="$envs FALN_MED3_$kv"; done
  f="gpurun_out/${TAG}_med_sweep_$(echo $cfg | tr ' =' '__').jsonl"
  echo "== sweep $cfg"; env $envs timeout 120 python tools/bench_med.py --quick > "$f" 2>&1; cut -c1-330 "$f"
done
echo "== full gpu suite"; timeout 900 python -m pytest tests -m gpu -q > gpurun_out/${TAG}_gputests.log 2>&1; echo "rc=$?"; tail -3 gpurun_out/${TAG}_gputests.log
echo "== bench"; timeout 400 python bench.py > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "rc=$?"; cut -c1-200 gpurun_out/${TAG}_bench.json
