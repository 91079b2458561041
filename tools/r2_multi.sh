#!/bin/bash
# usage: tools/r2_multi.sh N   (under gpurun --gpus N)
N=$1
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 200 --warmup 5 > gpurun_out/m_bench_n$N.json 2> gpurun_out/m_bench_n$N.err
cat gpurun_out/m_bench_n$N.json | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print(json.dumps({k:d[k] for k in ('value','n_gpus','ms_per_step','eager_ms_per_step','strong_scaling','clocks')}))
print(json.dumps({'e2e':d['e2e']['value'],'e2e_ms':d['e2e']['ms_per_step'],'ceiling':d['e2e']['host_copy_ceiling']}))
print(json.dumps(d['roofline']['kernel_ms_per_step']))
"
tail -3 gpurun_out/m_bench_n$N.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 200 --warmup 5 --no-graph > gpurun_out/m_bench_n${N}_eager.json 2> gpurun_out/m_bench_n${N}_eager.err
python -c "
import json
d=json.load(open('gpurun_out/m_bench_n${N}_eager.json'))
print('eager', d['value'], d['ms_per_step'])"
