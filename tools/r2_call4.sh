#!/bin/bash
mkdir -p gpurun_out
rm -f gpurun_out/reference_stack_parity.txt
timeout 900 python -m pytest tests/test_gpu_reference_stack.py -x -q -m gpu > gpurun_out/c4_pytest_stack.log 2>&1
echo "stack tests rc=$?" >> gpurun_out/c4_pytest_stack.log
tail -15 gpurun_out/c4_pytest_stack.log; cat gpurun_out/reference_stack_parity.txt gpurun_out/reference_test_py.txt 2>/dev/null
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "bench_configuration or large_channel or fused" > gpurun_out/c4_pytest_new.log 2>&1
echo "new tests rc=$?" >> gpurun_out/c4_pytest_new.log
tail -8 gpurun_out/c4_pytest_new.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/c4_smoke.log 2>&1; tail -3 gpurun_out/c4_smoke.log
timeout 400 python bench.py --steps 100 > gpurun_out/c4_bench.json 2> gpurun_out/c4_bench.err
cat gpurun_out/c4_bench.json; tail -5 gpurun_out/c4_bench.err
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/c4_bench_ref.json 2> gpurun_out/c4_bench_ref.err
cut -c1-300 gpurun_out/c4_bench_ref.json; tail -3 gpurun_out/c4_bench_ref.err
for args in "--amp" "--amp --fuse" "--amp --restated" "--amp --reference-module"; do timeout 200 python tools/encoder_bench.py $args 2>gpurun_out/c4_enc.err | tail -1; tail -2 gpurun_out/c4_enc.err; done > gpurun_out/c4_encoder_bench.jsonl
cat gpurun_out/c4_encoder_bench.jsonl
timeout 300 python tools/encoder_bench.py --amp --profile > gpurun_out/c4_enc_profile.json 2> gpurun_out/c4_enc_profile.txt
head -60 gpurun_out/c4_enc_profile.txt | cut -c1-200
