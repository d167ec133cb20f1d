#!/bin/bash
# Round-2 visit H (2 GPUs): single-GPU bench with every config, batch-1 launch list, then the 2-GPU bench + multi-GPU tests
set -u
mkdir -p gpurun_out
echo "== bench N=1"; ( time timeout 1500 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_full.json 2> gpurun_out/bench_full.err ) 2>&1 | tail -3; echo rc=$?; cut -c1-300 gpurun_out/bench_full.json; tail -5 gpurun_out/bench_full.err
echo "== batch-1 launch list"
CMD2="python tools/prof_case.py --rows 10000000 --batch 1 --iters 3"
ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/b1_launches.csv $CMD2 > gpurun_out/b1_launches_run.log 2>&1; tail -3 gpurun_out/b1_launches_run.log
echo "== pytest multi"; timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -q -x > gpurun_out/pytest_multi_2.log 2>&1; echo rc=$?; tail -4 gpurun_out/pytest_multi_2.log
echo "== bench N=2"; ( time timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err ) 2>&1 | tail -3; echo rc=$?; cut -c1-300 gpurun_out/bench_n2.json; tail -8 gpurun_out/bench_n2.err
