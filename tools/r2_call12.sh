#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "decoder or direct or gathers or gaps" > gpurun_out/c12_pytest.log 2>&1
echo "tests rc=$?" >> gpurun_out/c12_pytest.log
tail -3 gpurun_out/c12_pytest.log
timeout 600 python tools/sweep.py --decoder --scales 16 --points 4 --no-cpu > gpurun_out/c12_sweep.log 2>&1; tail -14 gpurun_out/c12_sweep.log | grep -v "^{"
