#!/bin/bash
# round-2: the driver's scaling sequence (N = 8 only, plus N = 1 on the same box for the same-box ratio)
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nvidia-smi -L | wc -l
N=${1:-8}
echo "== bench N=$N"; (time timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/r2n_bench_n$N.json 2> gpurun_out/r2n_bench_n$N.err); echo "rc=$?"; tail -2 gpurun_out/r2n_bench_n$N.err | cut -c1-200
echo "== bench N=1"; timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 --extras stage2,test --no-cpu-baseline > gpurun_out/r2n_bench_n1.json 2> gpurun_out/r2n_bench_n1.err; echo "rc=$?"
python - <<P
import json
for n in ($N, 1):
    r=json.loads(open(f'gpurun_out/r2n_bench_n{n}.json').read().strip().splitlines()[-1])
    print('N=%d stage1'%n, round(r['value'],1), round(r['ms_per_step'],3), 'e2e', round(r['e2e']['value'],1), r.get('comm'))
    for k,v in r.get('extras',{}).items():
        if 'value' in v: print('   extra',k,round(v['value'],1),round(v['ms_per_step'],3), v.get('comm'))
P
