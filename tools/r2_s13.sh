#!/bin/bash
set -u
cd "$(dirname "$0")/.."
run() { env "$@" timeout 300 python bench.py --steps 100 --no-extras --no-cpu-baseline 2>/dev/null | python -c "import sys,json; r=json.loads(sys.stdin.read()); print('   ', r['ms_per_step'])"; }
echo "== tests"; timeout 900 python -m pytest tests/test_model_gpu.py -m gpu -q --no-header 2>&1 | tail -2
for v in 1 0 1 0; do echo "BIAS_STREAM=$v"; run FALN_BIAS_STREAM=$v; done
echo "BIAS_STREAM=1 CHSUM_CAP=1"; run FALN_BIAS_STREAM=1 FALN_CHSUM_CAP=1
echo "BIAS_STREAM=1 CHSUM_CAP=4"; run FALN_BIAS_STREAM=1 FALN_CHSUM_CAP=4
echo "stage2 BIAS_STREAM=1"; FALN_BIAS_STREAM=1 timeout 600 python bench.py --workload stage2 --steps 40 --no-extras --no-cpu-baseline 2>/dev/null | python -c "import sys,json; r=json.loads(sys.stdin.read()); print('stage2', r['ms_per_step'])"
echo "stage2 BIAS_STREAM=0"; FALN_BIAS_STREAM=0 timeout 600 python bench.py --workload stage2 --steps 40 --no-extras --no-cpu-baseline 2>/dev/null | python -c "import sys,json; r=json.loads(sys.stdin.read()); print('stage2', r['ms_per_step'])"
