#!/bin/bash
mkdir -p gpurun_out
export MSDA_WALK_G4=1
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "golden or config1 or bit_reproducible or full_a2d or bench_configuration or frame_independent or odd_geom or bf16" > gpurun_out/c13_pytest.log 2>&1
echo "tests rc=$?" >> gpurun_out/c13_pytest.log
tail -4 gpurun_out/c13_pytest.log
timeout 200 python tools/level_breakdown.py --steps 20 2>/dev/null
for v in g4mb4 g4mb6; do MSDA_LIB=$PWD/variants/$v.so timeout 200 python tools/level_breakdown.py --steps 20 2>/dev/null | head -1; done
unset MSDA_WALK_G4
timeout 200 python tools/level_breakdown.py --steps 20 2>/dev/null | head -1
