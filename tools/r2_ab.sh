#!/bin/bash
# Round-2 A/B call: sub-bin target, scan unroll, rank sort v1/v2, unordered flag; new tests.
mkdir -p gpurun_out
cd $GRAFT_REPO_ROOT
python -m pytest tests/test_gpu_parity.py -q -x -k "unordered or bit_reproducible or oversized or golden or bench_configuration or frame_independent" > gpurun_out/ab_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/ab_pytest.log
tail -3 gpurun_out/ab_pytest.log
{
for lib in default rankv1 rankmb6 rankmb8 scan1 sbt3 sbt4 sbt8 sbt12; do
  if [ "$lib" = default ]; then unset MSDA_LIB; else export MSDA_LIB=$PWD/variants/$lib.so; fi
  for dt in bf16mix; do
    echo "=== $lib $dt"; python tools/kernel_times.py --dtype $dt --steps 30
  done
done
unset MSDA_LIB
echo "=== default fp32"; python tools/kernel_times.py --dtype fp32 --steps 30
echo "=== default N=2"; python tools/kernel_times.py --dtype bf16mix --steps 30 --N 2
echo "=== unordered"; python tools/kernel_times.py --dtype bf16mix --steps 30 --flags 128
echo "=== uniform? (probe)"; 
} > gpurun_out/ab_times.txt 2>&1
# the decoder one-launch backward under ncu (full set), 16 frames x 20 queries
ncu --set full --clock-control none -k regex:msda_grad_value_direct -c 2 --csv --page raw --log-file gpurun_out/full_direct.raw.csv python tools/kernel_times.py --lq 20 --dtype fp32 --steps 1 > gpurun_out/ncu_direct.log 2>&1
python tools/kernel_times.py --lq 20 --dtype fp32 --steps 30 >> gpurun_out/ab_times.txt 2>&1
python tools/kernel_times.py --lq 20 --dtype bf16mix --steps 30 >> gpurun_out/ab_times.txt 2>&1
cat gpurun_out/ab_times.txt | grep -v "^  memset\|big" | head -150
