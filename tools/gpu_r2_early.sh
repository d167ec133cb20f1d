#!/bin/bash
# early release of the accumulator stage (LK_DBG bit 32) against the default, interleaved, twice
set -u
run() { python tools/prof_case.py "$@" 2>&1 | tail -1 | sed 's/.*search kernel \([0-9.]*\) ms.*/\1/'; }
case_() { name=$1; shift; for rep in 1 2; do for dbg in 0 32; do echo "$name [dbg=$dbg] $(LK_DBG=$dbg run "$@")"; done; done; }
case_ 20M_b4096 --rows 20000000 --batch 4096 --iters 3
case_ 20M_b1024 --rows 20000000 --batch 1024 --iters 3
case_ 20M_b256 --rows 20000000 --batch 256 --iters 3
case_ 2Mx768_b4096_k100 --rows 2000000 --dim 768 --batch 4096 --k 100 --metric euclidean --iters 3
case_ d64 --rows 1000000 --dim 64 --batch 10000 --iters 3
case_ c1 --rows 20000 --batch 10000 --iters 4
