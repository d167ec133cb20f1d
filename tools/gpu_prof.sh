#!/bin/bash
# ncu evidence for the two regimes (run under gpurun, one GPU).  Each ncu command runs only
# after the same command line exited 0 without ncu.
set -u
mkdir -p gpurun_out
B4="python tools/prof_case.py --rows 4000000 --batch 4096 --iters 2"
B1="python tools/prof_case.py --rows 10000000 --batch 1 --iters 3"
$B4 > gpurun_out/prof_b4096_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:umma_search -s 1 -c 1 -f -o gpurun_out/prof_b4096 $B4 > gpurun_out/prof_b4096_ncu.log 2>&1
echo "b4096 rc=$?"; tail -3 gpurun_out/prof_b4096_plain.log
$B1 > gpurun_out/prof_b1_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:umma_search -s 1 -c 1 -f -o gpurun_out/prof_b1 $B1 > gpurun_out/prof_b1_ncu.log 2>&1
echo "b1 rc=$?"; tail -3 gpurun_out/prof_b1_plain.log
BENCH="python bench.py --rows 10000000 --steps 2 --warmup 3"
$BENCH > gpurun_out/bench_for_launches.json 2> gpurun_out/bench_for_launches.err && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv $BENCH > gpurun_out/launches_run.log 2>&1
echo "launches rc=$?"; tail -2 gpurun_out/launches_run.log; ls -la gpurun_out
