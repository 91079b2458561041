#!/bin/bash
mkdir -p gpurun_out; cd $GRAFT_REPO_ROOT
python -m pytest tests/test_gpu_parity.py -q -x -k "host_frame" 2>&1 | tail -3
for fpc in 1 2 4; do
  python bench.py --steps 30 --warmup 3 --no-cpu --no-ref-cuda --e2e-frames-per-chunk $fpc > gpurun_out/e2e_fpc$fpc.json 2> gpurun_out/e2e_fpc$fpc.err
  python - <<PY
import json
d=json.loads(open("gpurun_out/e2e_fpc$fpc.json").read().strip().splitlines()[-1])
e=d["e2e"]; print("fpc $fpc graph: e2e ms", round(e["ms_per_step"],3), "value", round(e["value"]/1e6,2), "ceiling ms", round(e["host_copy_ceiling"]["ms_per_step"],3), "| step ms", round(d["ms_per_step"],4), "unordered", d.get("unordered",{}) and round(d["unordered"]["ms_per_step"],4))
PY
done
python bench.py --steps 30 --warmup 3 --no-cpu --no-ref-cuda --no-graph > gpurun_out/e2e_nograph.json 2>/dev/null
python - <<PY
import json
d=json.loads(open("gpurun_out/e2e_nograph.json").read().strip().splitlines()[-1])
e=d["e2e"]; print("no graph: e2e ms", round(e["ms_per_step"],3), "| step ms", round(d["ms_per_step"],4))
PY
