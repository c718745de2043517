#!/bin/bash
set -u
cd "$(dirname "$0")/.."
run() { local e=$1; shift; env $e timeout 400 python bench.py --steps 60 --no-extras --no-cpu-baseline "$@" 2>/dev/null | python -c "import sys,json; r=json.loads(sys.stdin.read()); print('   ', r['ms_per_step'], r['e2e']['ms_per_step'])"; }
echo "== tests"; timeout 1500 python -m pytest tests -m gpu -q --no-header 2>&1 | tail -4 | cut -c1-300
for v in 0 1 0 1; do echo "stage1 STEM_FMA=$v"; run FALN_STEM_FMA=$v; done
for v in 0 1; do echo "stage2 STEM_FMA=$v"; run FALN_STEM_FMA=$v --workload stage2; echo "test STEM_FMA=$v"; run FALN_STEM_FMA=$v --workload test; done
