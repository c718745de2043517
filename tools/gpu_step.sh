#!/bin/bash
# gpurun session for step-level numbers: side-stream / channel_sum variants, full GPU suite.
set -u
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
TAG=${1:-t}
echo "== model tests"; timeout 600 python -m pytest tests/test_model_gpu.py tests/test_conv_gpu.py -q > gpurun_out/${TAG}_modeltests.log 2>&1; echo "rc=$?"; tail -3 gpurun_out/${TAG}_modeltests.log
for cfg in "2 2" "1 2" "3 2" "4 2" "2 1" "3 1"; do
  set -- $cfg
  echo "== bench stage1 side_streams=$1 chsum_cap=$2"
  FALN_SIDE_STREAMS=$1 FALN_CHSUM_CAP=$2 timeout 300 python bench.py > gpurun_out/${TAG}_bench_ss_$1_$2.json 2> gpurun_out/${TAG}_bench_ss_$1_$2.err
  python -c "import json;d=json.load(open('gpurun_out/${TAG}_bench_ss_$1_$2.json'));print(d['ms_per_step'],d['value'],d['e2e']['value'])"
done
for ss in 1 2 3; do
  echo "== stage2 side_streams=$ss"; FALN_SIDE_STREAMS=$ss timeout 400 python bench.py --workload stage2 > gpurun_out/${TAG}_bench_stage2_ss$ss.json 2> gpurun_out/${TAG}_bench_stage2_ss$ss.err
  python -c "import json;d=json.load(open('gpurun_out/${TAG}_bench_stage2_ss$ss.json'));print(d['ms_per_step'],d['value'],d['e2e']['value'])"
done
echo "== full gpu suite"; timeout 900 python -m pytest tests -m gpu -q > gpurun_out/${TAG}_gputests.log 2>&1; echo "rc=$?"; tail -3 gpurun_out/${TAG}_gputests.log
