#!/bin/bash
# build_variant.sh NAME "-DFOO=1 ..."  -> build/variants/libptb200_NAME.so (tuning builds; select with PTB200_LIB=...)
set -e
cd "$(dirname "$0")/.."
name=$1; defs=$2
out=build/variants; mkdir -p $out
PKG=raytracing-rust_b200
NV="-gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -fmad=false -Xcompiler -fPIC -Xptxas -v"
nvcc $NV $defs -c $PKG/csrc/wavefront.cu -o $out/wavefront_$name.o 2> $out/wavefront_$name.log
nvcc $NV $defs -c $PKG/csrc/lbvh_build.cu -o $out/lbvh_build_$name.o 2> $out/lbvh_build_$name.log
nvcc $NV $defs -c $PKG/csrc/cwbvh_build.cu -o $out/cwbvh_build_$name.o 2> $out/cwbvh_build_$name.log
nvcc $NV $defs -c $PKG/csrc/sah_build.cu -o $out/sah_build_$name.o 2> $out/sah_build_$name.log
nvcc $NV $defs -c $PKG/csrc/context.cu -o $out/context_$name.o 2> $out/context_$name.log
nvcc -shared -gencode arch=compute_100a,code=sm_100a -o $out/libptb200_$name.so $out/context_$name.o $out/lbvh_build_$name.o $out/cwbvh_build_$name.o $out/sah_build_$name.o $PKG/csrc/multi.o $out/wavefront_$name.o $PKG/host/ssml_loader.o $PKG/host/image_out.o $PKG/host/image_in.o -cudart static
grep -E "k_traceINS|k_shadeILi0" -A2 $out/wavefront_$name.log | grep -E "Used|spill" | tr '\n' ' '; echo
