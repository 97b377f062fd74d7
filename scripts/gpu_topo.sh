#!/bin/bash
N=${1:-4}
mkdir -p gpurun_out
export SPZB200_NO_REBUILD=1
{ nvidia-smi topo -m; lscpu | grep -i -E "numa|socket|model name|^CPU\(s\)"; ls /sys/devices/system/node/; for d in /sys/bus/pci/devices/*; do if [ "$(cat $d/vendor 2>/dev/null)" = "0x10de" ]; then echo "$d numa=$(cat $d/numa_node) class=$(cat $d/class)"; fi; done; which numactl; cat /sys/devices/system/node/node*/cpulist; cat /sys/devices/system/node/node*/meminfo | grep MemTotal; } > gpurun_out/topo_n${N}.txt 2>&1
cat gpurun_out/topo_n${N}.txt | head -60
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 3 > gpurun_out/bench_n${N}.json 2> gpurun_out/bench_n${N}.err; echo "bench N=$N rc=$?"
python -c "
import json,sys
d=json.loads([l for l in open('gpurun_out/bench_n${N}.json') if l.startswith('{')][-1])
print('N',d['n_gpus'],'value',d['value'],'enc',d['roofline']['encode']['achieved'],'dec',d['roofline']['decode']['achieved'],'e2e',d['e2e']['value'],d['e2e']['ms_per_step'])"
