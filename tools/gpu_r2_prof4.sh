#!/bin/bash
# ncu --set full of the search kernel on the 768-d top-100 case after the partially resident query tile
# (plain run first, then ONE kernel instance under ncu; one GPU)
set -u
mkdir -p gpurun_out
A="--rows 2000000 --dim 768 --batch 4096 --k 100 --metric euclidean --iters 2"
python tools/prof_case.py $A > gpurun_out/prof_k100q_plain.log 2>&1; echo plain rc=$?; tail -1 gpurun_out/prof_k100q_plain.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:umma_search -s 1 -c 1 -f -o gpurun_out/prof_k100q_r02 python tools/prof_case.py $A > gpurun_out/prof_k100q_ncu.log 2>&1; echo ncu rc=$?; tail -2 gpurun_out/prof_k100q_ncu.log
ls -la gpurun_out/prof_k100q_r02.ncu-rep
