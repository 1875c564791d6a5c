#!/usr/bin/env bash
# SpeedOfLight + memory sections for every kernel of one B=256 step (second iteration), CSV to gpurun_out/
mkdir -p gpurun_out
B=${1:-256}; MODEL=${2:-n}; TAG=${3:-r2}
timeout 300 python tools/ncu_target.py --model $MODEL --batch $B --iters 2 --nms 1 > gpurun_out/ncu_plain.log 2>&1 || { echo plain run failed; tail -5 gpurun_out/ncu_plain.log; exit 1; }
timeout 1500 ncu --section SpeedOfLight --section MemoryWorkloadAnalysis --section LaunchStats --section Occupancy \
  --clock-control none --launch-skip 115 -c 130 --csv --page raw --log-file gpurun_out/sections_${MODEL}_b${B}_${TAG}.csv \
  python tools/ncu_target.py --model $MODEL --batch $B --iters 2 --nms 1 > gpurun_out/ncu_run.log 2>&1
echo "ncu rc=$?"; tail -3 gpurun_out/ncu_run.log; wc -l gpurun_out/sections_${MODEL}_b${B}_${TAG}.csv
