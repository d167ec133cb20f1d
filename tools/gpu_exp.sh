#!/bin/bash
set -u
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "umma_kernel or large or sharding or ties or edge or mid" 2>&1 | tail -3
for R in 10000000 40000000; do
C="python tools/prof_case.py --rows $R --dim 384 --batch 4096 --k 10 --iters 4"
echo "== rows=$R rotate"; $C 2>&1 | tail -2
echo "== rows=$R no rotate"; LK_DBG=4 $C 2>&1 | tail -2
done
C="python tools/prof_case.py --rows 10000000 --dim 384 --batch 1024 --k 10 --iters 4"
echo "== B=1024 rotate"; $C 2>&1 | tail -1
echo "== B=1024 no rotate"; LK_DBG=4 $C 2>&1 | tail -1
