#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_encoder_layer.py -x -q -m gpu > gpurun_out/c7_pytest_layer.log 2>&1
echo "layer tests rc=$?" >> gpurun_out/c7_pytest_layer.log
tail -25 gpurun_out/c7_pytest_layer.log
for args in "--amp" "--amp --fuse" "--b200-layers"; do timeout 200 python tools/encoder_bench.py $args 2>gpurun_out/c7_enc.err | tail -1; grep -i "error\|Traceback" -A8 gpurun_out/c7_enc.err | head -20; done > gpurun_out/c7_encoder_bench.jsonl
cat gpurun_out/c7_encoder_bench.jsonl
timeout 300 python tools/encoder_bench.py --b200-layers --profile > gpurun_out/c7_enc_profile.json 2> gpurun_out/c7_enc_profile.txt
head -50 gpurun_out/c7_enc_profile.txt | cut -c1-100,150-215
