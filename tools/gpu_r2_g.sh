#!/bin/bash
set -u
mkdir -p gpurun_out
echo "== pytest ae"; timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "ae_ or latent or compressor" > gpurun_out/pytest_ae.log 2>&1; echo rc=$?; tail -12 gpurun_out/pytest_ae.log | cut -c1-220
echo "== prof ae bf16"; timeout 300 python tools/prof_ae.py --rows 1010000 --precision bf16 --iters 6 2>&1 | tail -3
echo "== waits"; LK_AE_PROF=1 timeout 300 python tools/prof_ae.py --rows 1010000 --precision bf16 --iters 1 2>&1 | tail -3 | cut -c1-300
