#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_encoder_layer.py -x -q -m gpu > gpurun_out/c8_pytest_layer.log 2>&1
echo "layer tests rc=$?" >> gpurun_out/c8_pytest_layer.log
tail -3 gpurun_out/c8_pytest_layer.log
for args in "--amp" "--b200-layers" "--b200-layers --graph" "--amp --graph"; do timeout 200 python tools/encoder_bench.py $args 2>gpurun_out/c8_enc.err | tail -1; grep -i "error\|Traceback" -A12 gpurun_out/c8_enc.err | tail -14; done > gpurun_out/c8_encoder_bench.jsonl
cat gpurun_out/c8_encoder_bench.jsonl
