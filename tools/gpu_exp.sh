#!/bin/bash
set -u
run() { echo "== $*"; python tools/prof_case.py "$@" --iters 4 2>&1 | tail -1; }
for B in 1 64 4096; do
run --rows 1250000 --dim 768 --batch $B --k 10 --metric euclidean
run --rows 1250000 --dim 768 --batch $B --k 100 --metric euclidean
done
run --rows 1250000 --dim 768 --batch 1024 --k 100 --metric euclidean
run --rows 1250000 --dim 768 --batch 1024 --k 10 --metric euclidean
