#!/bin/bash
# A/B aid: build libfb2_w4a16 with extra -D flags into csrc/build/variants/<name>.so
# (select it at run time with FB2_LIB=<path>).  usage: scripts/build_variant.sh name "-DFOO -DBAR=1"
set -e
cd "$(dirname "$0")/../ferrum-infer-rs_b200/csrc"
name=$1; flags=$2
out=build/variants/$name; mkdir -p $out
for f in api generic_kernels moe_kernels tc_gemm parallel checkpoint; do
  /usr/local/cuda/bin/nvcc $flags -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo \
    -Xcompiler -fPIC,-fvisibility=hidden -c $f.cu -o $out/$f.o &
done
wait
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -shared -o build/variants/$name.so $out/*.o -cudart shared -ldl
echo built build/variants/$name.so
