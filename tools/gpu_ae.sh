#!/bin/bash
# AE encoder visit: parity tests of both encoder kernels, then config 2 timings.
set -u
mkdir -p gpurun_out
echo "== pytest ae"
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "ae_ or latent" > gpurun_out/pytest_ae.log 2>&1 ; echo "rc=$?" ; tail -15 gpurun_out/pytest_ae.log
echo "== c2"
timeout 600 python tools/bench_configs.py --only c2 --out gpurun_out/configs_c2.jsonl > gpurun_out/configs_c2.log 2>&1 ; echo "rc=$?" ; tail -8 gpurun_out/configs_c2.log
