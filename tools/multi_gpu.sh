#!/bin/bash
# usage (under gpurun --gpus N): tools/multi_gpu.sh N [drivers]
#   bench.py on N ranks (weak scaling + the strong split it reports beside it), the tests that need two devices, and,
#   with "drivers", the whole-model DDP step, the encoder under DDP and the sharded inference pass.
#   Outputs: gpurun_out/m_*.json (tools/collect_evidence.py files them).
N=$1
mkdir -p gpurun_out
cd "${GRAFT_REPO_ROOT:-.}"
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 600 $TR --master-port 29511 bench.py --gpus $N --steps 200 --warmup 5 > gpurun_out/m_bench_n$N.json 2> gpurun_out/m_bench_n$N.err
cut -c1-400 gpurun_out/m_bench_n$N.json; grep -i "error\|Traceback" -A5 gpurun_out/m_bench_n$N.err | head -20
timeout 300 python -m pytest tests -m gpu -q -p no:cacheprovider -k "device_that_owns or frames" > gpurun_out/m_pytest_n$N.log 2>&1; tail -2 gpurun_out/m_pytest_n$N.log
if [ "$2" = "drivers" ]; then
  timeout 600 $TR --master-port 29521 tools/soc_step.py --steps 5 > gpurun_out/m_soc_n$N.json 2> gpurun_out/m_soc_n$N.err; cat gpurun_out/m_soc_n$N.json
  timeout 300 $TR --master-port 29523 tools/encoder_bench.py --amp 2> gpurun_out/m_encoder_n$N.err | tail -1 > gpurun_out/m_encoder_n$N.json; cat gpurun_out/m_encoder_n$N.json
  timeout 300 $TR --master-port 29524 tools/inference_bench.py --amp 2> gpurun_out/m_infer_n$N.err | tail -1 > gpurun_out/m_infer_n$N.json; cat gpurun_out/m_infer_n$N.json
fi
