#!/bin/bash
# Build libcasing_b200.so in-tree for sm_100a (B200).  No CPU fallback exists.
set -e
cd "$(dirname "$0")"
OUT=../libcasing_b200.so
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
FLAGS="-O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -Xcompiler -fPIC --expt-relaxed-constexpr"
for f in cs_kernels cs_vec cs_band cs_dd cs_dist cs_api; do
  if [ ! -f build/$f.o ] || [ $f.cu -nt build/$f.o ] || [ cs_common.cuh -nt build/$f.o ] || [ cs_internal.h -nt build/$f.o ] || [ cs_stencil.cuh -nt build/$f.o ] || [ cs_dist.h -nt build/$f.o ] || [ ../../include/casing_b200.h -nt build/$f.o ]; then
    mkdir -p build
    $NVCC $FLAGS ${PTXAS_V:+-Xptxas -v} -c $f.cu -o build/$f.o &
  fi
done
wait
$NVCC -shared -gencode arch=compute_100a,code=sm_100a -o $OUT build/cs_kernels.o build/cs_vec.o build/cs_band.o build/cs_dd.o build/cs_dist.o build/cs_api.o -lcudart -ldl
echo built $OUT
