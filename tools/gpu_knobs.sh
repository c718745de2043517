#!/bin/bash
set -u
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
TAG=${1:-n}
run() { local name=$1; shift; env "$@" timeout 200 python bench.py > gpurun_out/${TAG}_bench_$name.json 2> gpurun_out/${TAG}_bench_$name.err; python -c "import json;d=json.load(open('gpurun_out/${TAG}_bench_$name.json'));print('stage1 $name', d['ms_per_step'],d['value'],d['e2e']['value'])"; }
run deep100 FALN_CONV_DEEP_PCT=100
run deep400 FALN_CONV_DEEP_PCT=400
run narrow50 FALN_CONV_NARROW_PCT=50
run narrow200 FALN_CONV_NARROW_PCT=200
