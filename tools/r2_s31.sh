#!/bin/bash
# N = 2: tail bucket + per-bucket Adam, exposed communication
for cfg in "0 0" "1 1" "0 1"; do
  set -- $cfg
  echo "== FALN_BUCKET_ADAM=$1 FALN_TAIL_MB=$2"
  FALN_BUCKET_ADAM=$1 FALN_TAIL_MB=$2 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 100 --no-extras --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
r=json.loads([l for l in sys.stdin.read().strip().splitlines() if l.startswith('{')][-1])
print('stage1 N=2', round(r['ms_per_step'],4), 'value', round(r['value'],1), 'e2e', round(r['e2e']['ms_per_step'],4), r.get('comm'))
"
done
