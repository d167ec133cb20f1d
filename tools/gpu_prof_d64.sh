#!/bin/bash
set -u
mkdir -p gpurun_out
C="python tools/prof_case.py --rows 1000000 --dim 64 --batch 10000 --iters 3"
echo "== cg2"; $C 2>&1 | tail -1
echo "== cg1"; LK_CG=1 $C 2>&1 | tail -1
$C > gpurun_out/prof_d64_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:umma_search -s 1 -c 1 -f -o gpurun_out/prof_d64 $C > gpurun_out/prof_d64_ncu.log 2>&1
echo "rc=$?"
