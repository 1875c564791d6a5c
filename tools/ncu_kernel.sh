#!/usr/bin/env bash
# full ncu capture (with source) of launches of one kernel: tools/ncu_kernel.sh <tag> <kernel regex> <skip> <count> [batch] [model]
mkdir -p gpurun_out
TAG=$1; RE=$2; SKIP=$3; CNT=$4; B=${5:-256}; MODEL=${6:-n}
timeout 900 ncu --set full --import-source on --clock-control none -k regex:$RE --launch-skip $SKIP -c $CNT \
  -o gpurun_out/${TAG} -f python tools/ncu_target.py --model $MODEL --batch $B --iters 2 --nms 1 > gpurun_out/ncu_kernel.log 2>&1
echo "rc=$?"
ncu -i gpurun_out/${TAG}.ncu-rep --page raw --csv > gpurun_out/${TAG}_raw.csv 2>/dev/null
ncu -i gpurun_out/${TAG}.ncu-rep --page source --csv > gpurun_out/${TAG}_src.csv 2>/dev/null
ncu -i gpurun_out/${TAG}.ncu-rep --page details > gpurun_out/${TAG}_details.txt 2>/dev/null
ls -la gpurun_out/${TAG}*
