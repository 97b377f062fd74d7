#!/bin/bash
N=${1:-8}
mkdir -p gpurun_out
export SPZB200_NO_REBUILD=1
nvidia-smi -L > gpurun_out/multi${N}_box.txt; nproc >> gpurun_out/multi${N}_box.txt; free -g >> gpurun_out/multi${N}_box.txt
for n in $N; do
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $n --steps 20 --warmup 3 > gpurun_out/bench_n${n}.json 2> gpurun_out/bench_n${n}.err; echo "bench N=$n rc=$?"
python -c "
import json,sys
d=json.loads([l for l in open('gpurun_out/bench_n${n}.json') if l.startswith('{')][-1])
print('N',d['n_gpus'],'value',d['value'],'enc',d['roofline']['encode']['achieved'],'dec',d['roofline']['decode']['achieved'],'e2e',d['e2e']['value'],d['e2e']['ms_per_step'],d['clocks'])"
tail -2 gpurun_out/bench_n${n}.err
done
