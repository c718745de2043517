#!/bin/bash
# round-2 session 3: new coverage tests (variants, input pipeline, metrics), reference parity, full suite
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
rm -f gpurun_out/parity_r2.jsonl
echo "== new tests"; timeout 1200 python -m pytest tests/test_variants.py tests/test_input_pipeline.py tests/test_metrics.py tests/test_reference_gpu.py -m gpu -q --no-header > gpurun_out/r2s3_new.log 2>&1; echo "rc=$?"; tail -40 gpurun_out/r2s3_new.log | cut -c1-400
echo "== rest of gpu suite"; timeout 900 python -m pytest tests -m gpu -q --deselect tests/test_reference_gpu.py --deselect tests/test_metrics.py --deselect tests/test_variants.py --deselect tests/test_input_pipeline.py > gpurun_out/r2s3_gputests.log 2>&1; echo "rc=$?"; tail -5 gpurun_out/r2s3_gputests.log
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
