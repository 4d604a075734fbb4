#!/bin/bash
# Build a variant of libmirl_b200.so with train_fused.cu recompiled with extra flags:
#   profiles/build_fused_variant.sh NAME [-DFLAG=..]...   -> mirl_b200/_build/variants/libfused_NAME.so
# select with MIRL_B200_LIB=mirl_b200/_build/variants/libfused_NAME.so
set -e
cd "$(dirname "$0")/.."
name=$1; shift
mkdir -p mirl_b200/_build/variants
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC \
  --expt-relaxed-constexpr "$@" -c mirl_b200/csrc/train_fused.cu -o mirl_b200/_build/variants/train_fused_$name.o 2>/dev/null
objs=""
for f in c_api pe_simt pe_tc pe_train critic member_draw tc_gemm mlp_rows; do objs="$objs mirl_b200/_build/$f.o"; done
/usr/local/cuda/bin/nvcc -shared -o mirl_b200/_build/variants/libfused_$name.so $objs mirl_b200/_build/variants/train_fused_$name.o \
  -lcudart_static -lpthread -ldl -lrt 2>/dev/null
echo built mirl_b200/_build/variants/libfused_$name.so
