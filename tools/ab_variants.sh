#!/bin/bash
# A/B of kernel-variant builds (variants/*.so, tools/build_variant.sh): per-kernel times of one forward + backward
# usage: tools/ab_variants.sh <variant names or "default"> ...   (env DTYPES="bf16mix fp32")
mkdir -p gpurun_out
cd ${GRAFT_REPO_ROOT:-.}
for lib in "$@"; do
  if [ "$lib" = default ]; then unset MSDA_LIB; else export MSDA_LIB=$PWD/variants/$lib.so; fi
  for dt in ${DTYPES:-bf16mix fp32}; do
    echo "=== $lib $dt"; python tools/kernel_times.py --dtype $dt --steps 30 ${KT_ARGS:-}
  done
done 2>&1 | tee gpurun_out/ab_times.txt | grep -v "memset\|big"
