#!/bin/bash
set -u
mkdir -p gpurun_out
C="python tools/prof_case.py --rows 10000000 --dim 768 --batch 64 --k 100 --metric euclidean --iters 4"
echo "== seed on";  $C > gpurun_out/k100_seed.log 2>&1; tail -2 gpurun_out/k100_seed.log
echo "== seed off"; LK_SEED=0 $C > gpurun_out/k100_noseed.log 2>&1; tail -2 gpurun_out/k100_noseed.log
echo "== k10"; python tools/prof_case.py --rows 10000000 --dim 768 --batch 64 --k 10 --metric euclidean --iters 4 2>&1 | tail -1
echo "== k32"; python tools/prof_case.py --rows 10000000 --dim 768 --batch 64 --k 32 --metric euclidean --iters 4 2>&1 | tail -1
$C > /dev/null 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/k100_launches.csv $C > gpurun_out/k100_ncu.log 2>&1
echo "ncu rc=$?"
grep -v "^==" gpurun_out/k100_launches.csv | awk -F'","' '{print $5, $NF}' | tail -40
