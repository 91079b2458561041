#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "decoder or direct or gathers or gaps" > gpurun_out/c11_pytest.log 2>&1
echo "tests rc=$?" >> gpurun_out/c11_pytest.log
tail -12 gpurun_out/c11_pytest.log
timeout 600 python tools/sweep.py --decoder > gpurun_out/c11_sweep.log 2>&1; tail -40 gpurun_out/c11_sweep.log
