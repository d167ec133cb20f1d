#!/bin/bash
# Round-2 visit A (one GPU): parity tests incl. the BASELINE-scale ones, smoke, a scaled-down bench with every
# side configuration, then the full bench + the reference arm.
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
free -g > gpurun_out/host.txt; nproc >> gpurun_out/host.txt
echo "== pytest scale"; timeout 1500 python -m pytest tests/test_gpu_scale.py -m gpu -q -x -s > gpurun_out/pytest_scale.log 2>&1; echo rc=$?; tail -15 gpurun_out/pytest_scale.log
echo "== pytest all (minus scale)"; timeout 1500 python -m pytest tests -m gpu -q --ignore=tests/test_gpu_scale.py > gpurun_out/pytest_all.log 2>&1; echo rc=$?; tail -15 gpurun_out/pytest_all.log
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo rc=$?; tail -3 gpurun_out/smoke.log
echo "== bench small"; timeout 900 python bench.py --rows 10000000 --config-scale 0.1 --steps 3 --warmup 3 > gpurun_out/bench_small.json 2> gpurun_out/bench_small.err; rc=$?; echo rc=$rc; cut -c1-1500 gpurun_out/bench_small.json; tail -12 gpurun_out/bench_small.err
if [ $rc -eq 0 ]; then
  echo "== bench full"; ( time timeout 1500 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_full.json 2> gpurun_out/bench_full.err ) 2>&1 | tail -4; echo rc=$?; cut -c1-600 gpurun_out/bench_full.json; tail -5 gpurun_out/bench_full.err
  echo "== reference arm"; ( time timeout 900 python bench.py --impl reference --steps 5 --warmup 2 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err ) 2>&1 | tail -4; cut -c1-900 gpurun_out/bench_ref.json; tail -3 gpurun_out/bench_ref.err
fi
