#!/bin/bash
# partially resident query tile (LK_QRES_KB K blocks stay in shared memory) on the shapes whose tile does not fit
set -u
run() { python tools/prof_case.py "$@" 2>&1 | tail -1 | sed 's/.*search kernel \([0-9.]*\) ms.*/\1/'; }
for rep in 1 2; do
for kb in ${LK_QRES_LIST:-0 2 4 5 6}; do
  export LK_QRES_KB=$kb
  echo "qres_kb=$kb 2Mx768_b4096_k100 $(run --rows 2000000 --dim 768 --batch 4096 --k 100 --metric euclidean --iters 3)"
  echo "qres_kb=$kb 2Mx768_b4096_k10 $(run --rows 2000000 --dim 768 --batch 4096 --k 10 --metric euclidean --iters 3)"
  echo "qres_kb=$kb 2Mx768_b512_k100 $(run --rows 2000000 --dim 768 --batch 512 --k 100 --metric euclidean --iters 3)"
  echo "qres_kb=$kb fp32_20k_b10000 $(run --rows 20000 --batch 10000 --storage fp32 --iters 4)"
  echo "qres_kb=$kb fp32_2M_b4096 $(run --rows 2000000 --batch 4096 --storage fp32 --iters 3)"
  echo "qres_kb=$kb fp32_4M_b256 $(run --rows 4000000 --batch 256 --storage fp32 --iters 3)"
done; done
