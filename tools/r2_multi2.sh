#!/bin/bash
# usage: tools/r2_multi2.sh N   (under gpurun --gpus N): whole-model DDP step, encoder under DDP, sharded inference
N=$1
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 600 $TR --master-port 29521 tools/soc_step.py --steps 5 > gpurun_out/m2_soc_n$N.json 2> gpurun_out/m2_soc_n$N.err
cat gpurun_out/m2_soc_n$N.json; grep -i "error\|Traceback" -A5 gpurun_out/m2_soc_n$N.err | head -20
if [ "$N" = "8" ]; then
  timeout 600 $TR --master-port 29522 tools/soc_step.py --steps 5 --amp > gpurun_out/m2_soc_n${N}_amp.json 2> gpurun_out/m2_soc_n${N}_amp.err
  cat gpurun_out/m2_soc_n${N}_amp.json
  timeout 300 $TR --master-port 29523 tools/encoder_bench.py --amp > gpurun_out/m2_encoder_n$N.json 2> gpurun_out/m2_encoder_n$N.err
  cat gpurun_out/m2_encoder_n$N.json
  timeout 300 $TR --master-port 29524 tools/inference_bench.py --amp > gpurun_out/m2_infer_n$N.json 2> gpurun_out/m2_infer_n$N.err
  cat gpurun_out/m2_infer_n$N.json; grep -i "error\|Traceback" -A5 gpurun_out/m2_infer_n$N.err | head
  timeout 300 $TR --master-port 29525 tools/inference_bench.py --amp --graph > gpurun_out/m2_infer_graph_n$N.json 2> gpurun_out/m2_infer_graph_n$N.err
  cat gpurun_out/m2_infer_graph_n$N.json
  timeout 300 python tools/inference_bench.py --amp > gpurun_out/m2_infer_n1.json 2> gpurun_out/m2_infer_n1.err
  cat gpurun_out/m2_infer_n1.json
fi
