#!/bin/bash
# round-2 session 2: new metric / ms_pp tests, reference parity, full suite, the new bench line (N=1)
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
rm -f gpurun_out/parity_r2.jsonl
echo "== metrics + reference tests"; timeout 900 python -m pytest tests/test_metrics.py tests/test_reference_gpu.py -m gpu -q --no-header > gpurun_out/r2s2_new.log 2>&1; echo "rc=$?"; tail -30 gpurun_out/r2s2_new.log
echo "== full gpu suite"; timeout 900 python -m pytest tests -m gpu -q --deselect tests/test_reference_gpu.py --deselect tests/test_metrics.py > gpurun_out/r2s2_gputests.log 2>&1; echo "rc=$?"; tail -5 gpurun_out/r2s2_gputests.log
echo "== bench"; (time timeout 900 python bench.py > gpurun_out/r2s2_bench.json 2> gpurun_out/r2s2_bench.err); echo "rc=$?"; tail -5 gpurun_out/r2s2_bench.err; cut -c1-1500 gpurun_out/r2s2_bench.json
echo "== reference arm"; (time timeout 600 python bench.py --impl reference --steps 20 > gpurun_out/r2s2_refarm.json 2> gpurun_out/r2s2_refarm.err); echo "rc=$?"; cut -c1-600 gpurun_out/r2s2_refarm.json
nproc
