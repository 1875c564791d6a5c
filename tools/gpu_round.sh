#!/usr/bin/env bash
# tests + bench + per-op profile, each in its own process
mkdir -p gpurun_out
rm -f gpurun_out/summary.txt
timeout 900 python -m pytest tests/test_gpu_forward.py -m gpu -q -s -x > gpurun_out/pytest_fwd.log 2>&1
echo "pytest fwd exit $?" >> gpurun_out/summary.txt
timeout 600 python -m pytest tests/test_gpu_nms.py -m gpu -q > gpurun_out/pytest_nms.log 2>&1
echo "pytest nms exit $?" >> gpurun_out/summary.txt
timeout 600 python bench.py --steps 10 --warmup 3 --profile-json gpurun_out/profile_n256.json > gpurun_out/bench.log 2>&1
echo "bench exit $?" >> gpurun_out/summary.txt
cat gpurun_out/summary.txt
grep -E "max-abs|worst per-op|vs fp32|passed|failed|Error|error" gpurun_out/pytest_fwd.log | tail -40
tail -3 gpurun_out/pytest_nms.log
tail -2 gpurun_out/bench.log
