#!/bin/bash
set -u
for B in 64 4096; do
echo "== k10 B=$B"; python tools/prof_case.py --rows 10000000 --dim 384 --batch $B --k 10 --iters 3 2>&1 | tail -1
echo "== k32 B=$B"; python tools/prof_case.py --rows 10000000 --dim 384 --batch $B --k 32 --iters 3 2>&1 | tail -1
echo "== k100 B=$B"; python tools/prof_case.py --rows 10000000 --dim 384 --batch $B --k 100 --iters 3 2>&1 | tail -1
done
echo "== k100 B=64 noseed"; LK_SEED=0 python tools/prof_case.py --rows 10000000 --dim 384 --batch 64 --k 100 --iters 3 2>&1 | tail -1
echo "== k100 B=1024"; python tools/prof_case.py --rows 10000000 --dim 384 --batch 1024 --k 100 --iters 3 2>&1 | tail -1
echo "== k10 B=1024"; python tools/prof_case.py --rows 10000000 --dim 384 --batch 1024 --k 10 --iters 3 2>&1 | tail -1
