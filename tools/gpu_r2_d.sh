#!/bin/bash
# Round-2 visit D: fp32 storage on the tensor cores; AE pair kernel v2.
set -u
mkdir -p gpurun_out
echo "== pytest fp32 tc"; timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "fp32" > gpurun_out/pytest_fp32.log 2>&1; echo rc=$?; tail -15 gpurun_out/pytest_fp32.log
echo "== pytest ae"; timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "ae_ or latent or compressor" > gpurun_out/pytest_ae.log 2>&1; echo rc=$?; tail -5 gpurun_out/pytest_ae.log
echo "== prof ae bf16"; timeout 300 python tools/prof_ae.py --rows 1010000 --precision bf16 --iters 6 2>&1 | tail -4
CMD="python tools/prof_ae.py --rows 1010000 --precision bf16 --iters 2"
ncu --set full --clock-control none --import-source on -k regex:ae_pair -s 1 -c 1 -f -o gpurun_out/prof_ae_pair2 $CMD > gpurun_out/prof_ae_ncu.log 2>&1; echo "ncu rc=$?"
echo "== pytest everything else quick"; timeout 1200 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "not fp32 and not ae_" > gpurun_out/pytest_rest.log 2>&1; echo rc=$?; tail -5 gpurun_out/pytest_rest.log
