#!/bin/bash
set -u
mkdir -p gpurun_out
C="python tools/prof_case.py --rows 4000000 --batch 4096 --iters 2"
$C > gpurun_out/prof_cg2_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:umma_search -s 1 -c 1 -f -o gpurun_out/prof_cg2 $C > gpurun_out/prof_cg2_ncu.log 2>&1
echo "rc=$?"; tail -3 gpurun_out/prof_cg2_plain.log
