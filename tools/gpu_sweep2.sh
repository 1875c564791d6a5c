#!/usr/bin/env bash
# perf sweep over tuning env vars; each config "name:VAR=v VAR2=v" is one short bench run (no extras, no CPU legs)
mkdir -p gpurun_out
rm -f gpurun_out/sweep.txt
run() {
  name=$1; shift
  env "$@" timeout 300 python bench.py --steps ${STEPS:-10} --warmup 3 --no-cpu-baseline --no-extras --profile-json gpurun_out/profile_$name.json > gpurun_out/bench_$name.log 2>&1
  echo "$name rc=$? $(tail -1 gpurun_out/bench_$name.log | python -c 'import sys,json; d=json.loads(sys.stdin.readline()); print(d["value"], "img/s  fwd_ms", d["forward_ms_per_step"], "conv_ms", d["roofline"]["kernel_ms_per_step"], "frac", d["roofline"]["frac"], "e2e", d["e2e"]["value"])' 2>&1 | tail -1)" >> gpurun_out/sweep.txt
}
for cfg in "$@"; do
  name=${cfg%%:*}; vars=${cfg#*:}
  run $name $vars
done
cat gpurun_out/sweep.txt
