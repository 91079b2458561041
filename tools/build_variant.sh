#!/bin/bash
# Build a kernel-variant library for A/B runs: tools/build_variant.sh <name> [-DMACRO=.. ...]
# -> variants/<name>.so (git-ignored, shipped by gpurun; select with MSDA_LIB=$PWD/variants/<name>.so).
# -DMSDA_SLIM keeps only the bench shapes (D = 32, P = 4, fp32 and bf16 values with fp32 locations): ~4x faster build.
set -e
name=$1; shift
mkdir -p variants
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo --shared -Xcompiler -fPIC "$@" \
  -o variants/$name.so neurips2023_soc_b200/csrc/msda_api.cu
echo variants/$name.so
