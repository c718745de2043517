#!/bin/bash
# next-rows measurements
mkdir -p gpurun_out
timeout 600 python bench.py --steps 20 --extras next_rows --no-cpu-baseline > gpurun_out/r2s17_bench.json 2> gpurun_out/r2s17_bench.err
echo "exit $?"; tail -c 3000 gpurun_out/r2s17_bench.err
python - <<'P'
import json
r = json.loads(open('gpurun_out/r2s17_bench.json').read().strip().splitlines()[-1])
print(r['value'], r['ms_per_step'])
print(json.dumps(r['extras']['next_rows'], indent=1))
P
