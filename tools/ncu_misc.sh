#!/usr/bin/env bash
# full ncu captures (with source) of the non-conv kernels of one YOLO11n B=256 forward: tools/ncu_misc.sh <tag>
mkdir -p gpurun_out
TAG=${1:-misc}
for k in stem_mma attention sppf_pool dwconv3x3; do
  timeout 600 ncu --set full --import-source on --clock-control none -k regex:$k --launch-skip 1 -c 1 \
    -o gpurun_out/${TAG}_$k -f python tools/ncu_target.py --model n --batch 256 --iters 2 > gpurun_out/ncu_misc.log 2>&1
  echo "$k rc=$?"
  ncu -i gpurun_out/${TAG}_$k.ncu-rep --page details > gpurun_out/${TAG}_${k}_details.txt 2>/dev/null
done
