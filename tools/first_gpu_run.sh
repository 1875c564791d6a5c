#!/usr/bin/env bash
# First bring-up on a B200: diagnostics first (separate processes so a trapped kernel cannot poison
# the following steps), then the GPU tests, then a short bench.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
for impl in direct tc; do
  timeout 300 python tools/gpu_diag.py --impl $impl --size n --hw 64 > gpurun_out/diag_${impl}_n64.log 2>&1
  echo "diag $impl n64 exit $?" >> gpurun_out/summary.txt
done
timeout 300 python tools/gpu_diag.py --impl tc --size n --hw 160 > gpurun_out/diag_tc_n160.log 2>&1
echo "diag tc n160 exit $?" >> gpurun_out/summary.txt
timeout 300 python tools/gpu_diag.py --impl tc --size x --hw 64 > gpurun_out/diag_tc_x64.log 2>&1
echo "diag tc x64 exit $?" >> gpurun_out/summary.txt
timeout 900 python -m pytest tests/test_gpu_nms.py -m gpu -q > gpurun_out/pytest_nms.log 2>&1
echo "pytest nms exit $?" >> gpurun_out/summary.txt
timeout 1200 python -m pytest tests/test_gpu_forward.py -m gpu -q -s > gpurun_out/pytest_fwd.log 2>&1
echo "pytest fwd exit $?" >> gpurun_out/summary.txt
timeout 600 python bench.py --steps 5 --warmup 3 --profile-json gpurun_out/profile_n256.json > gpurun_out/bench.log 2>&1
echo "bench exit $?" >> gpurun_out/summary.txt
cat gpurun_out/summary.txt
tail -5 gpurun_out/diag_tc_n64.log
tail -3 gpurun_out/pytest_nms.log
tail -3 gpurun_out/pytest_fwd.log
tail -2 gpurun_out/bench.log
