#!/bin/bash
# One GPU call that regenerates what profiles/ holds for a round (run under gpurun; outputs in gpurun_out/):
#   parity suite, bench lines (ours + reference arm), ncu launch list of the bench command, ncu --set full of the
#   largest kernels, probe table (N = 16 and N = 2), config-5 sweep with decoder shapes, encoder-level, whole-model
#   and inference drivers.  tools/collect_evidence.py <tag> then files it under profiles/.
mkdir -p gpurun_out
cd "${GRAFT_REPO_ROOT:-.}"
timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -3 gpurun_out/pytest_gpu.log
timeout 300 python bench.py --impl reference > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err
timeout 400 python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err
cut -c1-300 gpurun_out/bench_n1.json
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:msda -c 400 --csv --log-file gpurun_out/launches_bench.csv python bench.py --steps 2 --warmup 1 --no-cpu --no-ref-cuda --no-graph > gpurun_out/launches_bench.log 2>&1
for dt in fp32 bf16mix; do
  timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:msda --csv --log-file gpurun_out/launches_$dt.csv python tools/one_step.py --dtype $dt --steps 2 > /dev/null 2>&1
done
timeout 400 ncu --set full --clock-control none --import-source on -k regex:"msda_(fwd_tile|bwd_sample_tile|bin_rank_sort|grad_value_walk)" -s 8 -c 4 -o gpurun_out/full_bf16mix python tools/one_step.py --dtype bf16mix --steps 3 > gpurun_out/ncu_full.log 2>&1
timeout 200 ncu --set full --clock-control none --import-source on -k regex:"msda_grad_value_direct" -c 1 -o gpurun_out/full_direct python tools/kernel_times.py --lq 20 --dtype fp32 --steps 1 > /dev/null 2>&1
# gpurun brings back at most 64 MiB: keep the raw-metric pages, not the reports
for r in full_bf16mix full_direct; do
  ncu -i gpurun_out/$r.ncu-rep --page raw --csv > gpurun_out/$r.raw.csv 2>/dev/null
  rm -f gpurun_out/$r.ncu-rep
done
timeout 400 python tools/probe.py --iters 15 --dists encoder,uniform > gpurun_out/probe_n16.log 2>&1; cp gpurun_out/probe.json gpurun_out/probe_n16.json
timeout 200 python tools/probe.py --iters 15 --dists encoder --N 2 > gpurun_out/probe_n2.log 2>&1; cp gpurun_out/probe.json gpurun_out/probe_n2.json
timeout 600 python tools/sweep.py --decoder > gpurun_out/sweep.log 2>&1; tail -34 gpurun_out/sweep.log
for args in "--amp" "--amp --fuse" "--b200-layers" "--b200-layers --graph"; do timeout 200 python tools/encoder_bench.py $args 2>/dev/null | tail -1; done > gpurun_out/encoder_bench.jsonl
for args in "--amp" "--amp --restated --graph"; do timeout 200 python tools/inference_bench.py $args 2>/dev/null | tail -1; done > gpurun_out/inference_bench.jsonl
timeout 300 python tools/soc_step.py --steps 5 2>/dev/null | tail -1 > gpurun_out/soc_step_n1.json
timeout 120 python tools/level_breakdown.py --steps 20 > gpurun_out/level_breakdown.jsonl 2>/dev/null
python tools/pcie_probe.py --trace > gpurun_out/host_pipeline_timeline.txt 2>&1
ls -la gpurun_out | head -50
