#!/bin/bash
mkdir -p gpurun_out
rm -f gpurun_out/reference_stack_parity.txt
timeout 1200 python -m pytest tests/test_gpu_reference_stack.py -q -m gpu > gpurun_out/c5_pytest_stack.log 2>&1
echo "stack tests rc=$?" >> gpurun_out/c5_pytest_stack.log
tail -6 gpurun_out/c5_pytest_stack.log; cat gpurun_out/reference_stack_parity.txt gpurun_out/reference_test_py.txt 2>/dev/null
for v in r1 cur r1 cur; do
  if [ $v = cur ]; then unset MSDA_LIB; else export MSDA_LIB=$PWD/variants/$v.so; fi
  timeout 200 python tools/level_breakdown.py --steps 20 2>/dev/null | head -1
done > gpurun_out/c5_walker_r1_vs_cur.jsonl
cat gpurun_out/c5_walker_r1_vs_cur.jsonl
