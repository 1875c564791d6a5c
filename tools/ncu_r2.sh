#!/usr/bin/env bash
# full ncu captures (with source) of the kernels VERDICT r1 names: fused depthwise (head.cls.0.3), stem, stride-2 gather (net.p2.0)
mkdir -p gpurun_out
CMD="python tools/ncu_target.py --model n --batch 256 --iters 2"
$CMD > gpurun_out/ncu_plain.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/ncu_plain.log; exit 1; }
cap() {  # tag regex skip count
  timeout 600 ncu --set full --import-source on --clock-control none -k "regex:$2" --launch-skip $3 -c $4 -o gpurun_out/$1 -f $CMD > gpurun_out/ncu_$1.log 2>&1
  echo "$1 rc=$?"
  ncu -i gpurun_out/$1.ncu-rep --page raw --csv > gpurun_out/$1_raw.csv 2>/dev/null
  ncu -i gpurun_out/$1.ncu-rep --page source --csv > gpurun_out/$1_src.csv 2>/dev/null
  ncu -i gpurun_out/$1.ncu-rep --page details > gpurun_out/$1_details.txt 2>/dev/null
}
for spec in "$@"; do
  IFS=: read tag re skip cnt <<< "$spec"
  cap $tag "$re" $skip $cnt
done
ls -la gpurun_out/*.ncu-rep
