#!/bin/bash
# round-2 final validation of the committed state: the driver's sequence (gpu tests, smoke, reference arm, bench N=1)
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
echo "== pytest -m gpu"; (time timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/r2z_gputests.log 2>&1); echo "rc=$?"; tail -2 gpurun_out/r2z_gputests.log | cut -c1-200
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1 | cut -c1-250
echo "== bench"; (time timeout 1200 python bench.py > gpurun_out/r2z_bench_n1.json 2> gpurun_out/r2z_bench_n1.err); echo "rc=$?"
python - <<'P'
import json
r=json.load(open('gpurun_out/r2z_bench_n1.json'))
print('stage1', round(r['value'],1), round(r['ms_per_step'],3), 'e2e', round(r['e2e']['value'],1), 'frac', round(r['roofline']['frac'],3), r['clocks'])
ex=r['extras']
for k in ('stage2','test'): print(k, round(ex[k]['value'],1), round(ex[k]['ms_per_step'],3))
cl=ex['conv_layers']; print('conv table sums ours/cudnn/bound', cl['sum_ours_us'], cl['sum_cudnn_bf16_us'], cl['sum_bound_us'])
for row in cl['rows']:
    extra=''
    if 'block_fwd' in row: extra=' | block fwd %.1f vs lib %.1f, dgrad %.1f'%(row['block_fwd']['ours_folded_us'],row['block_fwd']['cudnn_interpolate_conv_elu_us'],row['block_dgrad']['ours_folded_us'])
    print('%-22s fwd %6.1f/%6.1f  dgrad %6.1f/%6.1f  wgrad %6.1f/%6.1f  bound %5.1f%s'%(row['layer'],row['fwd']['ours_us'] or -1,row['fwd']['cudnn_bf16_us'],row['dgrad']['ours_us'] or -1,row['dgrad']['cudnn_bf16_us'],row['wgrad']['ours_us'] or -1,row['wgrad']['cudnn_bf16_us'],row['bound_us'],extra))
print(r['cpu_baseline'])
P
