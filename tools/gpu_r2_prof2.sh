#!/bin/bash
# ncu of the search kernel on the final build: tensor-bound (B=4096), HBM-bound (B=1), top-100 selector (B=4096, 768-d)
set -u
mkdir -p gpurun_out
B4="python tools/prof_case.py --rows 4000000 --batch 4096 --iters 2"
B1="python tools/prof_case.py --rows 10000000 --batch 1 --iters 3"
K100="python tools/prof_case.py --rows 2000000 --dim 768 --batch 4096 --k 100 --metric euclidean --iters 2"
$B4 > gpurun_out/prof_b4096_plain.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:umma_search -s 1 -c 1 -f -o gpurun_out/prof_b4096_r02 $B4 > gpurun_out/prof_b4096_ncu.log 2>&1; echo "b4096 rc=$?"; tail -1 gpurun_out/prof_b4096_plain.log
$B1 > gpurun_out/prof_b1_plain.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:umma_search -s 1 -c 1 -f -o gpurun_out/prof_b1_r02 $B1 > gpurun_out/prof_b1_ncu.log 2>&1; echo "b1 rc=$?"; tail -1 gpurun_out/prof_b1_plain.log
$K100 > gpurun_out/prof_k100_plain.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:umma_search -s 1 -c 1 -f -o gpurun_out/prof_k100_r02 $K100 > gpurun_out/prof_k100_ncu.log 2>&1; echo "k100 rc=$?"; tail -1 gpurun_out/prof_k100_plain.log
