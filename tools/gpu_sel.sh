#!/bin/bash
# selector / merge visit: parity tests of the search kernels, then config 4 timings (k = 10 and 100).
set -u
mkdir -p gpurun_out
echo "== pytest search"
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "not ae_ and not latent" > gpurun_out/pytest_sel.log 2>&1 ; echo "rc=$?" ; tail -15 gpurun_out/pytest_sel.log
echo "== c4"
timeout 900 python tools/bench_configs.py --only ${1:-c4} --out gpurun_out/configs_sel.jsonl > gpurun_out/configs_sel.log 2>&1 ; echo "rc=$?" ; tail -12 gpurun_out/configs_sel.log
