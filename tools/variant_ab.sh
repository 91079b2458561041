#!/bin/bash
# A/B of kernel-variant builds under variants/ (MSDA_LIB picks the library): launch lists per variant.
# usage: tools/variant_ab.sh "<lib or 'default'>:<dtype>" ...
mkdir -p gpurun_out
for spec in "$@"; do
  lib=${spec%%:*}; dt=${spec##*:}
  if [ "$lib" = default ]; then unset MSDA_LIB; else export MSDA_LIB=$PWD/variants/$lib; fi
  ncu --metrics gpu__time_duration.sum --clock-control none -k regex:msda --csv --log-file gpurun_out/ab_${lib}_$dt.csv python tools/one_step.py --dtype $dt --steps 2 > /dev/null 2>&1
  python - <<PY
import csv
rows=[r for r in csv.reader(open("gpurun_out/ab_${lib}_$dt.csv")) if len(r)>5]
hdr=rows[0]; ki=hdr.index("Kernel Name"); vi=hdr.index("Metric Value")
n=(len(rows)-1)//2
print("$lib $dt total_us %.1f  "%(sum(float(r[vi]) for r in rows[1+n:])/1e3) + "  ".join("%s=%.1f"%(r[ki].split("<")[0].replace("void ","").replace("msda_","")[:16], float(r[vi])/1e3) for r in rows[1+n:]))
PY
done
