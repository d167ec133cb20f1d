#!/bin/bash
set -u
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "umma_kernel or large or sharding or ties or edge or mid" 2>&1 | tail -3
run() { echo "== $*"; python tools/prof_case.py "$@" --iters 3 2>&1 | tail -1; }
run --rows 1000000 --dim 64 --batch 10000
run --rows 10000000 --dim 384 --batch 4096
run --rows 10000000 --dim 384 --batch 4096 --k 100
run --rows 10000000 --dim 384 --batch 1024
run --rows 10000000 --dim 384 --batch 64
run --rows 10000000 --dim 384 --batch 1
run --rows 10000000 --dim 768 --batch 4096 --k 100 --metric euclidean
