#!/bin/bash
mkdir -p gpurun_out; cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests -m gpu -q -x -p no:cacheprovider ${PYTEST_ARGS:-} > gpurun_out/check_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/check_pytest.log
tail -3 gpurun_out/check_pytest.log
bash tools/r2_ab.sh default ${VARIANTS:-} | grep "===\|walk\|us per"
KT_ARGS="--N 2" DTYPES=bf16mix bash tools/r2_ab.sh default | grep "===\|walk\|us per"
