#!/bin/bash
set -u
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "umma_kernel or large or sharding or ties or edge or mid" 2>&1 | tail -3
echo "== d64 b10000 1M"; python tools/prof_case.py --rows 1000000 --dim 64 --batch 10000 --iters 3 2>&1 | tail -1
echo "== d384 b4096 10M"; python tools/prof_case.py --rows 10000000 --dim 384 --batch 4096 --iters 4 2>&1 | tail -2
echo "== d384 b1024 10M"; python tools/prof_case.py --rows 10000000 --dim 384 --batch 1024 --iters 3 2>&1 | tail -1
echo "== d384 b64 10M"; python tools/prof_case.py --rows 10000000 --dim 384 --batch 64 --iters 3 2>&1 | tail -1
echo "== d384 b1 10M"; python tools/prof_case.py --rows 10000000 --dim 384 --batch 1 --iters 3 2>&1 | tail -1
