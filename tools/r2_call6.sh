#!/bin/bash
mkdir -p gpurun_out
timeout 600 python tools/soc_step.py --steps 5 > gpurun_out/c6_soc_n1.json 2> gpurun_out/c6_soc_n1.err
cat gpurun_out/c6_soc_n1.json; tail -5 gpurun_out/c6_soc_n1.err
timeout 600 python tools/soc_step.py --steps 5 --amp > gpurun_out/c6_soc_n1_amp.json 2> gpurun_out/c6_soc_n1_amp.err
cat gpurun_out/c6_soc_n1_amp.json; tail -3 gpurun_out/c6_soc_n1_amp.err
