#!/bin/bash
# round-2 session 5: row-kernel TMA-store epilogue + small ops: full parity, A/B timing, bench
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
echo "== full gpu suite"; timeout 1500 python -m pytest tests -m gpu -q -x --no-header > gpurun_out/r2s5_gputests.log 2>&1; echo "rc=$?"; tail -12 gpurun_out/r2s5_gputests.log | cut -c1-300
L="conv0_1.*,conv1_1.*,deconv1,iconv1+conv0 (folded),deconv2,conv1.0"
echo "== timing new"; timeout 300 python tools/conv_layers.py --time --iters 20 --layers "$L" --ops fwd,dgrad 2>&1 | tail -14
echo "== timing old (FALN_CONV_NO_TMA_OUT)"; FALN_CONV_NO_TMA_OUT=1 timeout 300 python tools/conv_layers.py --time --iters 20 --layers "$L" --ops fwd,dgrad 2>&1 | tail -14
echo "== bench stage1"; timeout 600 python bench.py --no-extras --no-cpu-baseline 2>/dev/null | python -c "import sys,json; r=json.loads(sys.stdin.read()); print('stage1', r['value'], r['ms_per_step'], r['e2e']['value'], r['gpu_launches'])"
echo "== bench stage2"; timeout 600 python bench.py --workload stage2 --no-extras --no-cpu-baseline 2>/dev/null | python -c "import sys,json; r=json.loads(sys.stdin.read()); print('stage2', r['value'], r['ms_per_step'], r['e2e']['value'], r['gpu_launches'])"
echo "== launch list"; timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/r2s5_launches_stage1.csv python tools/profile_step.py stage1 2 > /dev/null 2>&1; echo "rc=$?"
