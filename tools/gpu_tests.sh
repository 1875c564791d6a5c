#!/usr/bin/env bash
mkdir -p gpurun_out
rm -f gpurun_out/summary.txt
timeout 300 python tools/gpu_diag.py --impl tc --size n --hw 160 > gpurun_out/diag_tc_n160.log 2>&1
echo "diag tc n160 exit $?" >> gpurun_out/summary.txt
timeout 900 python -m pytest tests/test_gpu_nms.py -m gpu -q > gpurun_out/pytest_nms.log 2>&1
echo "pytest nms exit $?" >> gpurun_out/summary.txt
timeout 1200 python -m pytest tests/test_gpu_forward.py -m gpu -q -s > gpurun_out/pytest_fwd.log 2>&1
echo "pytest fwd exit $?" >> gpurun_out/summary.txt
cat gpurun_out/summary.txt
grep -E "max-abs|rel-rms|vs fp32|passed|failed" gpurun_out/pytest_fwd.log | tail -60
tail -3 gpurun_out/pytest_nms.log
