#!/bin/bash
# Build a variant of libmirl_b200.so with mlp_rows.cu recompiled with extra flags (-DMB_ROWS_BLOCK=48 -DMB_ROWS_STAGES=3 ...):
#   profiles/build_rows_variant.sh NAME [-DFLAG=..]...   -> mirl_b200/_build/variants/librows_NAME.so  (select with MIRL_B200_LIB=)
set -e
cd "$(dirname "$0")/.."
name=$1; shift
mkdir -p mirl_b200/_build/variants
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC \
  --expt-relaxed-constexpr -Xptxas -v "$@" -c mirl_b200/csrc/mlp_rows.cu -o mirl_b200/_build/variants/mlp_rows_$name.o 2>&1 | grep -E "registers|spill" | head -2
objs=""
for f in c_api pe_simt pe_tc pe_train critic member_draw tc_gemm train_fused; do objs="$objs mirl_b200/_build/$f.o"; done
/usr/local/cuda/bin/nvcc -shared -o mirl_b200/_build/variants/librows_$name.so $objs mirl_b200/_build/variants/mlp_rows_$name.o \
  -lcudart_static -lpthread -ldl -lrt 2>/dev/null
echo built mirl_b200/_build/variants/librows_$name.so
