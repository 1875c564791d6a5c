#!/usr/bin/env bash
# full ncu capture of one conv launch of YOLO11x (B=64): tools/ncu_x.sh <tag> <conv index>
mkdir -p gpurun_out
TAG=$1; IDX=$2
SKIP=$((164 + IDX))
timeout 900 ncu --set full --import-source on --clock-control none -k regex:conv_gemm --launch-skip $SKIP -c 1 \
  -o gpurun_out/${TAG} -f python tools/ncu_target.py --model x --batch 64 --iters 2 > gpurun_out/ncu_x.log 2>&1
echo "rc=$?"
ncu -i gpurun_out/${TAG}.ncu-rep --page raw --csv > gpurun_out/${TAG}_raw.csv 2>/dev/null
ncu -i gpurun_out/${TAG}.ncu-rep --page source --csv > gpurun_out/${TAG}_src.csv 2>/dev/null
