#!/bin/bash
mkdir -p gpurun_out; cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests -m gpu -q -x -p no:cacheprovider > gpurun_out/check_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/check_pytest.log
tail -3 gpurun_out/check_pytest.log
DTYPES=fp32 bash tools/r2_ab.sh default f32mb7 | grep "===\|walk\|us per"
python tools/kernel_times.py --lq 20 --dtype fp32 --steps 30
python tools/kernel_times.py --lq 20 --dtype bf16mix --steps 30
python tools/encoder_bench.py --b200-layers --steps 10 2>&1 | tail -2
python tools/encoder_bench.py --b200-layers --graph --steps 10 2>&1 | tail -2
