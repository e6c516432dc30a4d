#!/bin/bash
# A/B builds of the CUDA library: scripts/build_variant.sh NAME [-DMACRO ...]  ->  build/ab/libptl_NAME.so
# (run with PTL_LIB_PATH=build/ab/libptl_NAME.so; build/ is git-ignored but travels to the GPU box)
set -e
cd "$(dirname "$0")/.."
name=$1; shift
mkdir -p build/ab
python particulator.jl_b200/build.py -o build/ab/libptl_$name.so "$@" >/dev/null
echo built build/ab/libptl_$name.so
