#!/usr/bin/env bash
# full ncu capture (with source) of single conv launches: tools/ncu_one.sh <tag> <conv launch index in forward> [more indices]
mkdir -p gpurun_out
TAG=$1; shift
for IDX in "$@"; do
  SKIP=$((79 + IDX))
  timeout 600 ncu --set full --import-source on --clock-control none -k regex:conv_gemm --launch-skip $SKIP -c 1 \
    -o gpurun_out/${TAG}_conv${IDX} -f python tools/ncu_target.py --model n --batch 256 --iters 2 > gpurun_out/ncu_one.log 2>&1
  echo "idx $IDX rc=$?"
  ncu -i gpurun_out/${TAG}_conv${IDX}.ncu-rep --page raw --csv > gpurun_out/${TAG}_conv${IDX}_raw.csv 2>/dev/null
  ncu -i gpurun_out/${TAG}_conv${IDX}.ncu-rep --page source --csv > gpurun_out/${TAG}_conv${IDX}_src.csv 2>/dev/null
done
ls -la gpurun_out/${TAG}_*
