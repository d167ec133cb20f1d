#!/bin/bash
# Round-2 visit B: the CTA-pair AE encoder (parity + timing), the scale test that failed on the oracle side, TMEM read rate.
set -u
mkdir -p gpurun_out
echo "== pytest ae"; timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "ae_ or latent or compressor" > gpurun_out/pytest_ae.log 2>&1; echo rc=$?; tail -15 gpurun_out/pytest_ae.log
echo "== prof ae bf16"; timeout 300 python tools/prof_ae.py --rows 1010000 --precision bf16 --iters 6 2>&1 | tail -7
echo "== prof ae bf16 single-CTA"; LK_AE_PAIR=0 timeout 300 python tools/prof_ae.py --rows 1010000 --precision bf16 --iters 4 2>&1 | tail -4
echo "== prof ae fp32"; timeout 300 python tools/prof_ae.py --rows 1010000 --precision fp32 --iters 4 2>&1 | tail -4
echo "== pytest scale metrics"; timeout 900 python -m pytest tests/test_gpu_scale.py -m gpu -q -x -k "metrics or ties or maha" -s > gpurun_out/pytest_scale2.log 2>&1; echo rc=$?; grep -E "mahalanobis|passed|failed|Error" gpurun_out/pytest_scale2.log | tail -12
echo "== microbench (tmem)"; timeout 600 tools/_bin/microbench > gpurun_out/microbench.log 2>&1; tail -12 gpurun_out/microbench.log
