#!/bin/bash
# Final single-GPU confirmation of the committed tree: full GPU tests, smoke, bench, reference arm.
set -u
mkdir -p gpurun_out
echo "== pytest all"; timeout 1800 python -m pytest tests -m gpu -q > gpurun_out/pytest_all.log 2>&1; echo rc=$?; tail -3 gpurun_out/pytest_all.log | cut -c1-200
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo rc=$?; tail -2 gpurun_out/smoke.log
echo "== bench (driver's command line)"; ( time timeout 1500 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/bench_full.json 2> gpurun_out/bench_full.err ) 2>&1 | tail -3; echo rc=$?; cut -c1-200 gpurun_out/bench_full.json; tail -3 gpurun_out/bench_full.err
echo "== reference arm (driver's command line)"; ( time timeout 900 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err ) 2>&1 | tail -3; echo rc=$?; cut -c1-300 gpurun_out/bench_ref.json
