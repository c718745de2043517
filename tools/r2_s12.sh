#!/bin/bash
# round-2 session 12: re-tune the side-stream work (bias sums, weight-gradient split-K) against the faster main chain
set -u
cd "$(dirname "$0")/.."
run() { env "$@" timeout 300 python bench.py --steps 100 --no-extras --no-cpu-baseline 2>/dev/null | python -c "import sys,json; r=json.loads(sys.stdin.read()); print('   ', r['ms_per_step'])"; }
echo "base"; run FALN_X=0; run FALN_X=0
for c in 1 3 4 8; do echo "CHSUM_CAP=$c"; run FALN_CHSUM_CAP=$c; done
for f in 75 125 150 200; do echo "WGRAD_FILL_PCT=$f"; run FALN_WGRAD_FILL_PCT=$f; done
echo "SIDE_STREAMS=2"; run FALN_SIDE_STREAMS=2
echo "MAIN_PRIORITY=1"; run FALN_MAIN_PRIORITY=1
echo "CHSUM 4 + FILL 150"; run FALN_CHSUM_CAP=4 FALN_WGRAD_FILL_PCT=150
