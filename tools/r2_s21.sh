#!/bin/bash
# ncu evidence for the split-K cluster kernel and the cluster percentile kernel
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:splitk -c 8 -o /tmp/r2i_splitk -f python tools/conv_layers.py --layers "conv5_1.*,conv6_1.*,deconv6,conv5.0" --ops fwd,dgrad --iters 1 > gpurun_out/r2i_splitk.log 2>&1; echo "rc=$?"
ncu -i /tmp/r2i_splitk.ncu-rep --page raw --csv > gpurun_out/r2i_splitk.raw.csv 2>/dev/null
timeout 300 ncu --set full --clock-control none -k regex:percentile -c 2 -o /tmp/r2i_pct -f python - > gpurun_out/r2i_pct.log 2>&1 <<'P'
import torch
from fal_net_b200 import postproc
d = 120 * torch.rand(8, 1, 375, 1242, device="cuda")
for _ in range(2):
    postproc.percentile_rows(d, 95.0, add=1e-6)
torch.cuda.synchronize()
P
ncu -i /tmp/r2i_pct.ncu-rep --page raw --csv > gpurun_out/r2i_pct.raw.csv 2>/dev/null
timeout 600 python bench.py --steps 20 --extras next_rows,conv_layers --no-cpu-baseline > gpurun_out/r2i_bench.json 2>/dev/null
python - <<'P'
import json
r = json.loads(open('gpurun_out/r2i_bench.json').read().strip().splitlines()[-1])
print(r['value'], r['ms_per_step'], r['roofline']['frac'])
print(json.dumps(r['extras']['next_rows']['f1_ms_pp']))
t = r['extras']['conv_layers']
print({k: v for k, v in t.items() if k != 'rows'})
P
du -sh gpurun_out
