#!/bin/bash
set -u
mkdir -p gpurun_out
C="python tools/prof_case.py --rows 4000000 --dim 768 --batch 64 --k 100 --metric euclidean --iters 2"
$C > gpurun_out/prof_k100_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:umma_search -s 3 -c 1 -f -o gpurun_out/prof_k100 $C > gpurun_out/prof_k100_ncu.log 2>&1
echo "rc=$?"; tail -3 gpurun_out/prof_k100_plain.log
C="python tools/prof_case.py --rows 4000000 --dim 768 --batch 64 --k 10 --metric euclidean --iters 2"
$C > gpurun_out/prof_k10_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:umma_search -s 1 -c 1 -f -o gpurun_out/prof_k10 $C > gpurun_out/prof_k10_ncu.log 2>&1
echo "rc=$?"; tail -3 gpurun_out/prof_k10_plain.log
