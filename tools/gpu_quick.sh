#!/usr/bin/env bash
# quick correctness + perf check after a kernel change: parity + forward tests, then a short bench with per-op profile
mkdir -p gpurun_out
rm -f gpurun_out/summary.txt
for f in parity forward fused_filter; do
  timeout 420 python -m pytest tests/test_gpu_$f.py -m gpu -q -s -x > gpurun_out/pytest_$f.log 2>&1
  echo "pytest $f exit $? $(tail -n 1 gpurun_out/pytest_$f.log)" >> gpurun_out/summary.txt
done
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-extras --profile-json gpurun_out/profile_new.json > gpurun_out/bench_new.log 2>&1
echo "bench exit $? $(tail -n 1 gpurun_out/bench_new.log | python -c 'import sys,json; d=json.loads(sys.stdin.readline()); print(d["value"], "img/s  fwd_ms", d["forward_ms_per_step"], "conv_ms", d["roofline"]["kernel_ms_per_step"], "frac", d["roofline"]["frac"], "e2e", d["e2e"]["value"])' 2>&1 | tail -n 1)" >> gpurun_out/summary.txt
cat gpurun_out/summary.txt
grep -hE "widehead|x@640|n@1280|bench tensor|e2e n@640|Error|assert" gpurun_out/pytest_parity.log | tail -n 20
grep -hE "worst per-op|Error|assert|max err" gpurun_out/pytest_forward.log | tail -n 30
