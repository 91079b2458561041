#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_encoder_layer.py tests/test_gpu_parity.py -x -q -m gpu -k "layer or raw or fused or encoder" > gpurun_out/c9_pytest.log 2>&1
echo "tests rc=$?" >> gpurun_out/c9_pytest.log
tail -12 gpurun_out/c9_pytest.log
for args in "--b200-layers" "--b200-layers --graph"; do timeout 200 python tools/encoder_bench.py $args 2>gpurun_out/c9_enc.err | tail -1; grep -i "error\|Traceback" -A12 gpurun_out/c9_enc.err | tail -14; done > gpurun_out/c9_encoder_bench.jsonl
cat gpurun_out/c9_encoder_bench.jsonl
timeout 300 python tools/encoder_bench.py --b200-layers --profile > /dev/null 2> gpurun_out/c9_enc_profile.txt
head -34 gpurun_out/c9_enc_profile.txt | cut -c1-100,150-215
