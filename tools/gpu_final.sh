#!/bin/bash
# Validation session: full GPU suite, smoke(), the four bench workloads, launch list of the Stage-1 step.
set -u
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
TAG=${1:-f}
echo "== full gpu suite"; timeout 900 python -m pytest tests -m gpu -q > gpurun_out/${TAG}_gputests.log 2>&1; echo "rc=$?"; tail -3 gpurun_out/${TAG}_gputests.log
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${TAG}_smoke.log 2>&1; echo "rc=$?"; tail -2 gpurun_out/${TAG}_smoke.log | cut -c1-300
for wl in stage1 stage2 test med; do
  echo "== bench $wl"; timeout 400 python bench.py --workload $wl > gpurun_out/${TAG}_bench_$wl.json 2> gpurun_out/${TAG}_bench_$wl.err; echo "rc=$?"
  python -c "import json;d=json.load(open('gpurun_out/${TAG}_bench_$wl.json'));print(d['metric'],d['ms_per_step'],d['value'],d['e2e']['value'])"
done
echo "== bench_med"; timeout 300 python tools/bench_med.py > gpurun_out/${TAG}_med_full.jsonl 2> gpurun_out/${TAG}_med.err; echo "rc=$?"
echo "== launch list stage1"
timeout 300 ncu --metrics gpu__time_duration.sum,launch__grid_size,launch__block_size --clock-control none -c 3000 --csv \
  --log-file gpurun_out/${TAG}_launches_stage1.csv python tools/profile_step.py stage1 2 > gpurun_out/${TAG}_launches.log 2>&1; echo "rc=$?"
echo "== reference arm"; timeout 400 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${TAG}_bench_reference.json 2> gpurun_out/${TAG}_bench_reference.err; echo "rc=$?"; cut -c1-300 gpurun_out/${TAG}_bench_reference.json
