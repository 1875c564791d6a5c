#!/usr/bin/env bash
# round 2: all GPU tests (each file in its own process) + the default bench line
mkdir -p gpurun_out
rm -f gpurun_out/summary.txt
for f in parity forward fused_filter nms preprocess metric; do
  timeout 420 python -m pytest tests/test_gpu_$f.py -m gpu -q -s > gpurun_out/pytest_$f.log 2>&1
  echo "pytest $f exit $?" >> gpurun_out/summary.txt
done
timeout 500 python bench.py --steps 20 --warmup 5 --profile-json gpurun_out/profile_n256.json > gpurun_out/bench.log 2> gpurun_out/bench.err
echo "bench exit $?" >> gpurun_out/summary.txt
cat gpurun_out/summary.txt
grep -hE "widehead|sweep|x@640|n@1280|bench tensor|e2e n@640|passed|failed|Error" gpurun_out/pytest_parity.log | tail -40
grep -hE "worst per-op|passed|failed|Error" gpurun_out/pytest_forward.log | tail -30
tail -2 gpurun_out/pytest_fused_filter.log gpurun_out/pytest_nms.log gpurun_out/pytest_preprocess.log gpurun_out/pytest_metric.log
tail -c 6000 gpurun_out/bench.log
tail -5 gpurun_out/bench.err
