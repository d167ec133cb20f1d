#!/bin/bash
# One GPU-box visit: diagnostics, parity tests (process-isolated groups), small bench.
# Usage (from the repo root, under gpurun):  bash tools/gpu_round.sh [rows]
set -u
mkdir -p gpurun_out
ROWS=${1:-10000000}
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
echo "== diag" ; timeout 300 python tools/umma_diag.py > gpurun_out/diag.log 2>&1 ; echo "diag rc=$?" ; tail -25 gpurun_out/diag.log
echo "== pytest non-umma"
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "not umma and not large and not sharding and not bf16_mid and not edge and not ties and not metrics and not faiss and not latent and not mahalanobis_bf16" > gpurun_out/pytest_a.log 2>&1 ; echo "rc=$?" ; tail -8 gpurun_out/pytest_a.log
echo "== pytest all"
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/pytest_all.log 2>&1 ; echo "rc=$?" ; tail -30 gpurun_out/pytest_all.log
echo "== smoke"
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1 ; echo "rc=$?" ; tail -3 gpurun_out/smoke.log
echo "== bench rows=$ROWS"
timeout 900 python bench.py --rows $ROWS --steps 3 --warmup 3 > gpurun_out/bench_small.json 2> gpurun_out/bench_small.err ; echo "rc=$?" ; cat gpurun_out/bench_small.json ; tail -5 gpurun_out/bench_small.err
