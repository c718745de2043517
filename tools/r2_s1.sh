#!/bin/bash
# round-2 session 1: reference-on-CUDA parity tests (verbose), then the full gpu suite
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
rm -f gpurun_out/parity_r2.jsonl
echo "== reference-on-CUDA parity"; timeout 900 python -m pytest tests/test_reference_gpu.py -m gpu -q -x --no-header -rA > gpurun_out/r2s1_ref.log 2>&1; echo "rc=$?"; tail -40 gpurun_out/r2s1_ref.log
timeout 900 python -m pytest tests/test_reference_gpu.py -m gpu -q --no-header > gpurun_out/r2s1_ref_all.log 2>&1; echo "rc(all)=$?"; tail -15 gpurun_out/r2s1_ref_all.log
echo "== full gpu suite"; timeout 900 python -m pytest tests -m gpu -q --deselect tests/test_reference_gpu.py > gpurun_out/r2s1_gputests.log 2>&1; echo "rc=$?"; tail -3 gpurun_out/r2s1_gputests.log
cat gpurun_out/parity_r2.jsonl | cut -c1-600
