#!/bin/bash
# A/B builds of the CUDA library: scripts/build_variant.sh NAME [-DMACRO ...]  ->  build/ab/libptl_NAME.so
# (run with PTL_LIB_PATH=build/ab/libptl_NAME.so; build/ is git-ignored but travels to the GPU box)
set -e
cd "$(dirname "$0")/.."
name=$1; shift
mkdir -p build/ab
cd particulator.jl_b200/csrc
env -u CC -u CXX nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -Xcompiler -fPIC,-fvisibility=hidden -shared --expt-relaxed-constexpr "$@" -o ../../build/ab/libptl_$name.so ptl_api.cu
echo built build/ab/libptl_$name.so
