#!/bin/bash
# tools/build_variant.sh NAME [-DFLAG ...]: builds quasi_b200/lib/variants/libquasi_cuda_NAME.so with extra flags
name=$1; shift
mkdir -p quasi_b200/lib/variants
exec /usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++20 -lineinfo -fmad=false -prec-div=true \
  -prec-sqrt=true -ftz=false --cudart shared -Xcompiler -fPIC,-fvisibility=hidden,-ffp-contract=off -Iinclude -Iquasi_b200/csrc \
  "$@" -shared -o quasi_b200/lib/variants/libquasi_cuda_$name.so quasi_b200/csrc/plugin.cu quasi_b200/csrc/pathtrace.cu \
  quasi_b200/csrc/mesh_trace.cu quasi_b200/csrc/mesh_wavefront.cu quasi_b200/csrc/lbvh_build.cu quasi_b200/csrc/post.cu
