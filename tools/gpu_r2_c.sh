#!/bin/bash
# Round-2 visit C: ncu of the CTA-pair AE encoder; stage-count sensitivity; remaining scale tests.
set -u
mkdir -p gpurun_out
CMD="python tools/prof_ae.py --rows 1010000 --precision bf16 --iters 2"
$CMD > gpurun_out/prof_ae_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:ae_pair -s 1 -c 1 -f -o gpurun_out/prof_ae_pair $CMD > gpurun_out/prof_ae_ncu.log 2>&1
echo "ncu rc=$?"; tail -2 gpurun_out/prof_ae_plain.log
for s in 2 4 6; do echo "stages $s"; LK_AE_STAGES=$s python tools/prof_ae.py --rows 1010000 --precision bf16 --iters 4 2>&1 | tail -1; done
echo "== pytest scale rest"; timeout 900 python -m pytest tests/test_gpu_scale.py -m gpu -q -k "ties or maha" -s > gpurun_out/pytest_scale3.log 2>&1; echo rc=$?; grep -E "mahalanobis|passed|failed|Error" gpurun_out/pytest_scale3.log | tail -12
