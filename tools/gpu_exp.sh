#!/bin/bash
set -u
for R in 10000000 25000000 50000000 100000000; do
echo "== rows=$R B=1"; python tools/prof_case.py --rows $R --dim 384 --batch 1 --iters 6 2>&1 | tail -3
done
