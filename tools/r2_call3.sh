#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "tile_kernel or bit_reproducible or many_samples or config1 or odd_geom or gaps" > gpurun_out/c3_pytest_new.log 2>&1
echo "new tests rc=$?" >> gpurun_out/c3_pytest_new.log
tail -5 gpurun_out/c3_pytest_new.log
timeout 200 python tools/level_breakdown.py --steps 10 > gpurun_out/c3_levels.jsonl 2> gpurun_out/c3_levels.err
cat gpurun_out/c3_levels.jsonl; tail -3 gpurun_out/c3_levels.err
timeout 200 python tools/level_breakdown.py --steps 10 --dtype fp32 > gpurun_out/c3_levels_fp32.jsonl 2> gpurun_out/c3_levels_fp32.err
head -1 gpurun_out/c3_levels_fp32.jsonl
timeout 200 python tools/level_breakdown.py --steps 10 --dist uniform > gpurun_out/c3_levels_uniform.jsonl 2> gpurun_out/c3_levels_uniform.err
head -1 gpurun_out/c3_levels_uniform.jsonl
timeout 200 python tools/level_breakdown.py --steps 10 --N 2 > gpurun_out/c3_levels_n2.jsonl 2> gpurun_out/c3_levels_n2.err
head -1 gpurun_out/c3_levels_n2.jsonl
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"msda_bwd_bin" -s 2 -c 1 -o gpurun_out/c3_bin python tools/one_step.py --dtype bf16mix --steps 3 > gpurun_out/c3_ncu.log 2>&1
ncu -i gpurun_out/c3_bin.ncu-rep --page raw --csv > gpurun_out/c3_bin.raw.csv 2>/dev/null
ncu -i gpurun_out/c3_bin.ncu-rep --page source --csv > gpurun_out/c3_bin.src.csv 2>/dev/null
rm -f gpurun_out/c3_bin.ncu-rep
python tools/ncu_summary.py gpurun_out/c3_bin.raw.csv | head -24
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/c3_pytest_all.log 2>&1
echo "all tests rc=$?" >> gpurun_out/c3_pytest_all.log
tail -4 gpurun_out/c3_pytest_all.log
