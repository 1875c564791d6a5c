#!/usr/bin/env bash
# Round evidence: (1) plain bench, (2) ncu launch list of the same command, (3) DRAM bytes + duration of every
# conv launch of one B=256 forward, (4) full captures of a few dominant launches.  TAG=r01 by default.
TAG=${1:-r02}
mkdir -p gpurun_out
timeout 600 python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-extras > gpurun_out/${TAG}_bench_plain.log 2>&1 || { echo "plain bench failed"; tail -3 gpurun_out/${TAG}_bench_plain.log; exit 1; }
timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_launches.csv \
  python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-extras > gpurun_out/${TAG}_ncu_bench.log 2>&1
echo "launch list rc=$?"
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_bytes.sum,sm__pipe_tc_cycles_active.avg.pct_of_peak_sustained_active,sm__cycles_elapsed.max \
  --clock-control none -k regex:conv_gemm --launch-skip 79 -c 79 --csv --log-file gpurun_out/${TAG}_conv_dram.csv \
  python tools/ncu_target.py --model n --batch 256 --iters 2 > gpurun_out/${TAG}_ncu_dram.log 2>&1
echo "conv dram rc=$?"
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum \
  --clock-control none --launch-skip 85 -c 85 --csv --log-file gpurun_out/${TAG}_all_dram.csv \
  python tools/ncu_target.py --model n --batch 256 --iters 2 --nms 1 > gpurun_out/${TAG}_ncu_all.log 2>&1
echo "all dram rc=$?"
bash tools/ncu_one.sh ${TAG} 0 3 61 64 66
