#!/bin/bash
mkdir -p gpurun_out; cd $GRAFT_REPO_ROOT
timeout 400 ncu --set full --clock-control none --import-source on -k regex:"msda_grad_value_walk" -s 2 -c 1 -o gpurun_out/walk python tools/one_step.py --dtype bf16mix --steps 3 > gpurun_out/ncu_walk.log 2>&1
ncu -i gpurun_out/walk.ncu-rep --page raw --csv > gpurun_out/walk.raw.csv 2>/dev/null
ncu -i gpurun_out/walk.ncu-rep --page source --csv > gpurun_out/walk.source.csv 2>/dev/null
rm -f gpurun_out/walk.ncu-rep
ls -la gpurun_out/walk.*
