#!/bin/bash
set -u
mkdir -p gpurun_out
echo "== pytest exchange"; timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "exchange or sharding or merge" > gpurun_out/pytest_x.log 2>&1; echo rc=$?; tail -4 gpurun_out/pytest_x.log | cut -c1-200
echo "== pytest multi"; timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -q -x > gpurun_out/pytest_multi_2.log 2>&1; echo rc=$?; tail -3 gpurun_out/pytest_multi_2.log
echo "== bench N=2"; timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err; echo rc=$?; cut -c1-200 gpurun_out/bench_n2.json; grep -v "OMP_NUM\|^\*\*\*\|^$" gpurun_out/bench_n2.err | tail -4
