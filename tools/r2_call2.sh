#!/bin/bash
mkdir -p gpurun_out
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"msda_grad_value_tile" -s 2 -c 1 -o gpurun_out/c2_tile python tools/one_step.py --dtype bf16mix --steps 3 > gpurun_out/c2_ncu.log 2>&1
ncu -i gpurun_out/c2_tile.ncu-rep --page raw --csv > gpurun_out/c2_tile.raw.csv 2>/dev/null
ncu -i gpurun_out/c2_tile.ncu-rep --page source --csv > gpurun_out/c2_tile.src.csv 2>/dev/null
rm -f gpurun_out/c2_tile.ncu-rep
python tools/ncu_summary.py gpurun_out/c2_tile.raw.csv
for v in step8; do
  MSDA_LIB=$PWD/variants/$v.so timeout 200 python tools/level_breakdown.py --steps 10 > gpurun_out/c2_levels_$v.jsonl 2> gpurun_out/c2_levels_$v.err
  head -1 gpurun_out/c2_levels_$v.jsonl; tail -2 gpurun_out/c2_levels_$v.err
done
