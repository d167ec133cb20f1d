#!/bin/bash
# Full-size bench (100M x 384) + reference arm + ncu evidence.  Run under gpurun, one GPU.
set -u
mkdir -p gpurun_out
echo "== reference arm"; timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo rc=$?; cut -c1-400 gpurun_out/bench_ref.json
echo "== full bench"; timeout 1500 python bench.py > gpurun_out/bench_full.json 2> gpurun_out/bench_full.err; echo rc=$?; cat gpurun_out/bench_full.json; tail -3 gpurun_out/bench_full.err
bash tools/gpu_prof.sh
