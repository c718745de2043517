#!/bin/bash
# round-2 session 8 (2 GPUs): the driver's N=2 launch of bench.py (headline + extras), reference arm under torchrun
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nvidia-smi -L
echo "== bench N=2"; (time timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r2s8_bench_n2.json 2> gpurun_out/r2s8_bench_n2.err); echo "rc=$?"; tail -3 gpurun_out/r2s8_bench_n2.err; python - <<'P'
import json
r=json.loads(open('gpurun_out/r2s8_bench_n2.json').read().strip().splitlines()[-1])
print('N=2 stage1', r['value'], r['ms_per_step'], r['e2e']['value'])
for k,v in r.get('extras',{}).items(): print('  extra',k,v.get('value'),v.get('ms_per_step'))
print(r['roofline']['frac'], r['roofline'].get('frac_of_layerwise_roofline'), r['roofline'].get('frac_of_layerwise_roofline_ss'))
P
echo "== bench N=1"; (time timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r2s8_bench_n1.json 2> gpurun_out/r2s8_bench_n1.err); echo "rc=$?"; python - <<'P'
import json
r=json.loads(open('gpurun_out/r2s8_bench_n1.json').read().strip().splitlines()[-1])
print('N=1 stage1', r['value'], r['ms_per_step'], r['e2e']['value'])
for k,v in r.get('extras',{}).items(): print('  extra',k,v.get('value'),v.get('ms_per_step'))
print(r['roofline']['frac'], r['roofline'].get('frac_of_layerwise_roofline'), r['roofline'].get('frac_of_layerwise_roofline_ss'))
P
echo "== reference arm N=2"; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29518 bench.py --impl reference --gpus 2 --steps 20 --warmup 5 2>/dev/null | tail -1 | cut -c1-300
