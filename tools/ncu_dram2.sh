# DRAM bytes of conv launches [skip, skip+count) of the second forward: tools/ncu_dram2.sh <skip> <count> [env...]
SKIP=$1; CNT=$2; shift 2
env "$@" ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,lts__t_sectors_srcunit_tex_op_read.sum,lts__t_sectors_srcunit_tex_op_read_lookup_miss.sum --clock-control none -k regex:conv_gemm --launch-skip $((79 + SKIP)) -c $CNT --csv python tools/ncu_target.py --model n --batch 256 --iters 2 2>/dev/null | grep -E "dram__bytes|gpu__time|lts__" | awk -F'","' '{print $5, $(NF-2), $(NF-1), $NF}'
