#!/bin/bash
# gpurun session for step-level numbers: stream priority / footprint variants of the gradient side work.
set -u
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
TAG=${1:-t}
run() {  # name, env...
  local name=$1; shift
  echo "== bench stage1 $name"
  env "$@" timeout 300 python bench.py > gpurun_out/${TAG}_bench_$name.json 2> gpurun_out/${TAG}_bench_$name.err
  python -c "import json;d=json.load(open('gpurun_out/${TAG}_bench_$name.json'));print(d['ms_per_step'],d['value'],d['e2e']['value'])"
}
run p100 FALN_WGRAD_FILL_PCT=100
run p75 FALN_WGRAD_FILL_PCT=75
run p50 FALN_WGRAD_FILL_PCT=50
run p33 FALN_WGRAD_FILL_PCT=33
run p50_cs1 FALN_WGRAD_FILL_PCT=50 FALN_CHSUM_CAP=1
run p200 FALN_WGRAD_FILL_PCT=200
for pct in 100 50; do
echo "== stage2 pct=$pct"; FALN_WGRAD_FILL_PCT=$pct timeout 400 python bench.py --workload stage2 > gpurun_out/${TAG}_bench_stage2_p$pct.json 2> gpurun_out/${TAG}_bench_stage2_p$pct.err
python -c "import json;d=json.load(open('gpurun_out/${TAG}_bench_stage2_p$pct.json'));print(d['ms_per_step'],d['value'],d['e2e']['value'])"
done
