#!/bin/bash
# round-2 final evidence session: ncu --set full of the MED kernels at the six config-#5 shapes and of the conv kernels on
# representative layers (raw pages exported to CSV on the box; the .ncu-rep files exceed gpurun's 64 MiB return limit),
# launch lists of a Stage-1 and a Stage-2 step, MED microbench, the full default bench line (N=1)
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
echo "== ncu MED six shapes"
for shp in 8,33,375,1242 8,49,375,1242 8,65,375,1242 2,33,1024,2048 2,49,1024,2048 2,65,1024,2048; do
  tag=$(echo $shp | tr ',' 'x')
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:"med3_|med_disp" -c 6 -o /tmp/r2f_med_$tag -f python tools/bench_med.py --profile $shp > gpurun_out/r2f_med_$tag.log 2>&1; echo "$shp rc=$?"
  ncu -i /tmp/r2f_med_$tag.ncu-rep --page raw --csv > gpurun_out/r2f_med_$tag.raw.csv 2>/dev/null
done
echo "== ncu conv"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:conv3x3 -c 16 -o /tmp/r2f_conv -f python tools/conv_layers.py --layers "deconv1,conv0_1.*,iconv3,conv5_1.*,conv3_1.*" --ops fwd,dgrad,wgrad --iters 1 > gpurun_out/r2f_conv.log 2>&1; echo "rc=$?"
ncu -i /tmp/r2f_conv.ncu-rep --page raw --csv > gpurun_out/r2f_conv.raw.csv 2>/dev/null
echo "== launch lists"
timeout 900 ncu --metrics gpu__time_duration.sum,launch__grid_size,launch__block_size --clock-control none -c 3000 --csv --log-file gpurun_out/r2f_launches_stage1.csv python tools/profile_step.py stage1 2 > /dev/null 2>&1; echo "rc=$?"
timeout 900 ncu --metrics gpu__time_duration.sum,launch__grid_size,launch__block_size --clock-control none -c 6000 --csv --log-file gpurun_out/r2f_launches_stage2.csv python tools/profile_step.py stage2 2 > /dev/null 2>&1; echo "rc=$?"
echo "== med microbench"; timeout 400 python tools/bench_med.py > gpurun_out/r2f_bench_med.jsonl 2>&1; cut -c1-200 gpurun_out/r2f_bench_med.jsonl
echo "== full bench"; (time timeout 1200 python bench.py > gpurun_out/r2f_bench_n1.json 2> gpurun_out/r2f_bench_n1.err); echo "rc=$?"; cut -c1-300 gpurun_out/r2f_bench_n1.json
du -sh gpurun_out; ls -la gpurun_out/r2f_* | head -30
