#!/bin/bash
# Build an experimental variant of libmirl_b200.so: profiles/build_variant.sh NAME [-DFLAG=..]...
# Only pe_tc.cu is recompiled with the extra flags; select with MIRL_B200_LIB=mirl_b200/_build/variants/lib_NAME.so
set -e
cd "$(dirname "$0")/.."
name=$1; shift
mkdir -p mirl_b200/_build/variants
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC \
  --expt-relaxed-constexpr "$@" -c mirl_b200/csrc/pe_tc.cu -o mirl_b200/_build/variants/pe_tc_$name.o 2>/dev/null
objs=""
for f in c_api pe_simt pe_train critic member_draw tc_gemm train_fused mlp_rows; do objs="$objs mirl_b200/_build/$f.o"; done
/usr/local/cuda/bin/nvcc -shared -o mirl_b200/_build/variants/lib_$name.so $objs mirl_b200/_build/variants/pe_tc_$name.o \
  -lcudart_static -lpthread -ldl -lrt 2>/dev/null
echo built mirl_b200/_build/variants/lib_$name.so
