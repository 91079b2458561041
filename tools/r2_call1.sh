#!/bin/bash
# round 2, GPU call 1: correctness of the grad_value tile kernel, then timings against the first-generation pair
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/c1_gpu.txt
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "tile_kernel or bit_reproducible or many_samples or config1" > gpurun_out/c1_pytest_new.log 2>&1
echo "new tests rc=$?" >> gpurun_out/c1_pytest_new.log
tail -5 gpurun_out/c1_pytest_new.log
timeout 200 python tools/level_breakdown.py --steps 10 > gpurun_out/c1_levels_tile.jsonl 2> gpurun_out/c1_levels_tile.err
timeout 200 python tools/level_breakdown.py --steps 10 --flags 32 > gpurun_out/c1_levels_v1.jsonl 2> gpurun_out/c1_levels_v1.err
timeout 200 python tools/level_breakdown.py --steps 10 --dtype fp32 > gpurun_out/c1_levels_tile_fp32.jsonl 2> gpurun_out/c1_levels_tile_fp32.err
timeout 200 python tools/level_breakdown.py --steps 10 --dist uniform > gpurun_out/c1_levels_tile_uniform.jsonl 2> gpurun_out/c1_levels_tile_uniform.err
timeout 200 python tools/level_breakdown.py --steps 10 --N 2 > gpurun_out/c1_levels_tile_n2.jsonl 2> gpurun_out/c1_levels_tile_n2.err
cat gpurun_out/c1_levels_tile.jsonl gpurun_out/c1_levels_v1.jsonl
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/c1_pytest_all.log 2>&1
echo "all tests rc=$?" >> gpurun_out/c1_pytest_all.log
tail -5 gpurun_out/c1_pytest_all.log
timeout 300 python bench.py --steps 100 --no-cpu > gpurun_out/c1_bench.json 2> gpurun_out/c1_bench.err
cat gpurun_out/c1_bench.json
