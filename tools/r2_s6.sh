#!/bin/bash
# round-2 session 6: parity after the small-op kernels; stage-1/2 bench; MED microbench (win6 shuffle, streaming disp); launch list
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
echo "== full gpu suite"; timeout 1500 python -m pytest tests -m gpu -q --no-header > gpurun_out/r2s6_gputests.log 2>&1; echo "rc=$?"; tail -12 gpurun_out/r2s6_gputests.log | cut -c1-300
echo "== bench stage1"; timeout 600 python bench.py --no-extras --no-cpu-baseline 2>/dev/null | python -c "import sys,json; r=json.loads(sys.stdin.read()); print('stage1', r['value'], r['ms_per_step'], r['e2e']['value'], r['gpu_launches'])"
echo "== bench stage2"; timeout 600 python bench.py --workload stage2 --no-extras --no-cpu-baseline 2>/dev/null | python -c "import sys,json; r=json.loads(sys.stdin.read()); print('stage2', r['value'], r['ms_per_step'], r['e2e']['value'], r['gpu_launches'])"
echo "== med microbench"; timeout 300 python tools/bench_med.py --quick 2>&1 | cut -c1-420
echo "== launch list"; timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/r2s6_launches_stage1.csv python tools/profile_step.py stage1 2 > /dev/null 2>&1; echo "rc=$?"
