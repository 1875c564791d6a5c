#!/usr/bin/env bash
# Builds libyolob200.so (sm_100a only) and the oracle's C restatement, in-tree.
set -euo pipefail
cd "$(dirname "$0")"
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
OUT=yolo_infer_pt_b200/lib
mkdir -p "$OUT" build
SRCS="api plan conv_tc kernels_misc nms preprocess metric"
OBJS=""
for s in $SRCS; do
  src=yolo_infer_pt_b200/csrc/$s.cu
  obj=build/$s.o
  if [ ! -f "$obj" ] || [ "$src" -nt "$obj" ] || [ yolo_infer_pt_b200/csrc/yb_internal.h -nt "$obj" ] || [ include/yolob200.h -nt "$obj" ]; then
    "$NVCC" -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 \
      -Xcompiler -fPIC -Xptxas -v ${NVCC_EXTRA:-} -c "$src" -o "$obj" 2> "build/$s.ptxas.log" || { cat "build/$s.ptxas.log"; exit 1; }
  fi
  OBJS="$OBJS $obj"
done
"$NVCC" -shared -gencode arch=compute_100a,code=sm_100a -o "$OUT/libyolob200.so" $OBJS -lcuda
gcc -O2 -fPIC -shared -fno-fast-math -ffp-contract=off -o oracle/libnms_oracle.so oracle/nms_oracle.c -lm
echo "built $OUT/libyolob200.so and oracle/libnms_oracle.so"
