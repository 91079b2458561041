#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu --deselect tests/test_gpu_reference_stack.py::test_reference_ops_test_script_passes_on_this_extension > gpurun_out/c10_pytest_all.log 2>&1
echo "all tests rc=$?" >> gpurun_out/c10_pytest_all.log
tail -4 gpurun_out/c10_pytest_all.log
timeout 200 python tools/level_breakdown.py --steps 20 2>/dev/null | head -1
timeout 200 python tools/level_breakdown.py --steps 20 --N 2 2>/dev/null | head -1
timeout 300 python bench.py --steps 200 --no-cpu > gpurun_out/c10_bench.json 2> gpurun_out/c10_bench.err
python -c "
import json
d=json.load(open('gpurun_out/c10_bench.json'))
print(d['value'], d['ms_per_step'], d['eager_ms_per_step'], d['roofline']['kernel_ms_per_step'], d['roofline']['step']['frac'])"
