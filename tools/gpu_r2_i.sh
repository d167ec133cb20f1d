#!/bin/bash
set -u
mkdir -p gpurun_out
echo "== pytest umma"; timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "umma or ties or mid or edge or scratch or sharding or large or fp32 or deep or merge" > gpurun_out/pytest_umma.log 2>&1; echo rc=$?; tail -3 gpurun_out/pytest_umma.log | cut -c1-200
echo "== pytest scale"; timeout 900 python -m pytest tests/test_gpu_scale.py -m gpu -q -x -k "not maha" > gpurun_out/pytest_scale.log 2>&1; echo rc=$?; tail -2 gpurun_out/pytest_scale.log | cut -c1-200
echo "== c1"; python tools/prof_case.py --rows 20000 --batch 10000 --iters 4 2>&1 | tail -1
echo "== d64"; python tools/prof_case.py --rows 1000000 --dim 64 --batch 10000 --iters 3 2>&1 | tail -1
echo "== 10M b4096"; python tools/prof_case.py --rows 10000000 --batch 4096 --iters 3 2>&1 | tail -1
echo "== 10M b256"; python tools/prof_case.py --rows 10000000 --batch 256 --iters 3 2>&1 | tail -1
echo "== 10M b64"; python tools/prof_case.py --rows 10000000 --batch 64 --iters 3 2>&1 | tail -1
echo "== 10M b1"; python tools/prof_case.py --rows 10000000 --batch 1 --iters 3 2>&1 | tail -1
echo "== 2M x768 b4096 k100"; python tools/prof_case.py --rows 2000000 --dim 768 --batch 4096 --k 100 --metric euclidean --iters 2 | tail -1
echo "== 10M x768 b64 k100"; python tools/prof_case.py --rows 10000000 --dim 768 --batch 64 --k 100 --metric euclidean --iters 3 | tail -1
