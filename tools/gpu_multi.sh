#!/bin/bash
# Multi-GPU visit: NCCL sharded-search test + bench at N GPUs.  bash tools/gpu_multi.sh N
set -u
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv > gpurun_out/gpus_$N.txt
echo "== pytest multi"; timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -q -x > gpurun_out/pytest_multi_$N.log 2>&1; echo rc=$?; tail -5 gpurun_out/pytest_multi_$N.log
echo "== bench N=$N"
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err; echo rc=$?; cat gpurun_out/bench_n$N.json; tail -5 gpurun_out/bench_n$N.err
