#!/bin/bash
# A/B of search-kernel build variants parked under tools/_bin/ (tools/build_variant.sh), interleaved per case and
# repeated (the box's clocks drift with temperature over a run: variants measured back to back, twice)
set -u
mkdir -p gpurun_out
run() { python tools/ab_old_lib.py "$@" 2>&1 | tail -1 | sed 's/.*search kernel \([0-9.]*\) ms.*/\1/'; }
V=${LK_AB_VARIANTS:-base share pack}
case_() { name=$1; shift; for rep in 1 2; do for v in $V; do echo "$name [$v] $(run tools/_bin/lib_$v.so "$@")"; done; done; }
case_ c1 --rows 20000 --batch 10000 --iters 4
case_ d64 --rows 1000000 --dim 64 --batch 10000 --iters 3
case_ 4M_b4096 --rows 4000000 --batch 4096 --iters 3
case_ 10M_b256 --rows 10000000 --batch 256 --iters 3
case_ 2Mx768_b4096_k100 --rows 2000000 --dim 768 --batch 4096 --k 100 --metric euclidean --iters 2
case_ 10Mx768_b64_k100 --rows 10000000 --dim 768 --batch 64 --k 100 --metric euclidean --iters 3
