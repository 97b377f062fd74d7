#!/bin/bash
# usage: gpu_multi.sh N   -- bench on N GPUs via torchrun + the multi-GPU host entry point test
N=${1:-2}
mkdir -p gpurun_out
export SPZB200_NO_REBUILD=1
nvidia-smi -L > gpurun_out/multi_box.txt; nproc >> gpurun_out/multi_box.txt; free -g >> gpurun_out/multi_box.txt
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 3 > gpurun_out/bench_n${N}.json 2> gpurun_out/bench_n${N}.err; echo "bench N=$N rc=$?"
cat gpurun_out/bench_n${N}.json | cut -c1-1500
tail -3 gpurun_out/bench_n${N}.err
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "multi or context" 2>&1 | tail -3
timeout 600 python bench.py --impl reference --gpus $N --steps 3 --warmup 1 > gpurun_out/bench_ref_n${N}.json 2>&1; echo "ref rc=$?"; cut -c1-600 gpurun_out/bench_ref_n${N}.json
