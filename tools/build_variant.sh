#!/bin/bash
# Build a kernel-variant library for A/B runs: tools/build_variant.sh <name> [-DMACRO=.. ...]
# -> variants/<name>.so (git-ignored, shipped by gpurun; select with MSDA_LIB=$PWD/variants/<name>.so).
set -e
name=$1; shift
python -m neurips2023_soc_b200.build --variant "$name" "$@"
