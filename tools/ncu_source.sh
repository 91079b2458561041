#!/bin/bash
# source-level ncu pages (per-SASS-line executed instructions and stall samples) of the kernels named in $1 (regex)
mkdir -p gpurun_out; cd $GRAFT_REPO_ROOT
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"$1" -s ${SKIP:-4} -c ${COUNT:-2} -o gpurun_out/src python tools/one_step.py --dtype ${DT:-bf16mix} --steps 3 > gpurun_out/ncu_src.log 2>&1
ncu -i gpurun_out/src.ncu-rep --page raw --csv > gpurun_out/src.raw.csv 2>/dev/null
for k in $(echo "$1" | tr '|' ' ' | tr -d '()'); do
  ncu -i gpurun_out/src.ncu-rep --page source --csv --kernel-name regex:$k > gpurun_out/src_$k.source.csv 2>/dev/null
done
rm -f gpurun_out/src.ncu-rep
ls -la gpurun_out/src*
