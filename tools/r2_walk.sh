#!/bin/bash
# walker with warp-convergent loops: parity suite + per-kernel times
mkdir -p gpurun_out
cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests -m gpu -q -x -p no:cacheprovider > gpurun_out/walk_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/walk_pytest.log
tail -4 gpurun_out/walk_pytest.log
{
for dt in bf16mix fp32 bf16; do echo "=== N=16 $dt"; python tools/kernel_times.py --dtype $dt --steps 30; done
echo "=== N=2"; python tools/kernel_times.py --dtype bf16mix --steps 30 --N 2
echo "=== N=8"; python tools/kernel_times.py --dtype bf16mix --steps 30 --N 8
python tools/level_breakdown.py 2>&1 | tail -12
python tools/probe.py --iters 10 --dists uniform 2>&1 | tail -8
} > gpurun_out/walk_times.txt 2>&1
grep -v "memset\|big" gpurun_out/walk_times.txt | head -90
