#!/bin/bash
set -u
mkdir -p gpurun_out
for n in 8; do
echo "== bench N=$n"; ( time timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2951$n bench.py --gpus $n --steps 10 --warmup 3 > gpurun_out/bench_n$n.json 2> gpurun_out/bench_n$n.err ) 2>&1 | tail -3; echo rc=$?; cut -c1-250 gpurun_out/bench_n$n.json; grep -v "OMP_NUM\|^\*\*\*\|^$" gpurun_out/bench_n$n.err | tail -5
done
