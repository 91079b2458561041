#!/bin/bash
# One GPU development cycle: parity suite, timing probe, per-kernel launch list.
mkdir -p gpurun_out
timeout 1000 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -${TAILN:-6} gpurun_out/pytest_gpu.log
timeout 400 python tools/probe.py --iters 10 --dists ${DISTS:-encoder} ${PROBE_ARGS:-} > gpurun_out/probe.log 2>&1
tail -4 gpurun_out/probe.log
for dt in ${DTYPES:-fp32 bf16mix}; do
  ncu --metrics gpu__time_duration.sum --clock-control none -k regex:msda --csv --log-file gpurun_out/launches_$dt.csv python tools/one_step.py --dtype $dt --steps 2 > gpurun_out/ncu_$dt.log 2>&1
done
python - <<PY
import csv
for dt in "${DTYPES:-fp32 bf16mix}".split():
    rows=[r for r in csv.reader(open(f"gpurun_out/launches_{dt}.csv")) if len(r)>5]
    hdr=rows[0]; ki=hdr.index("Kernel Name"); vi=hdr.index("Metric Value")
    n=(len(rows)-1)//2
    print(dt, "total_us", sum(float(r[vi]) for r in rows[1+n:])/1e3)
    for r in rows[1+n:]: print("  %-66s %10s"%(r[ki][:66], r[vi]))
PY
