#!/bin/bash
set -u
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "umma_kernel or large or sharding" 2>&1 | tail -3
for D in 384 768; do
for B in 256 4096; do
C="python tools/prof_case.py --rows 10000000 --dim $D --batch $B --k 10 --iters 4"
echo "== D=$D B=$B cg=2"; $C 2>&1 | tail -1
echo "== D=$D B=$B cg=1"; LK_CG=1 $C 2>&1 | tail -1
done
done
