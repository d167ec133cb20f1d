#!/bin/bash
set -u
mkdir -p gpurun_out
echo "== pytest ae"; timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "ae_ or latent or compressor" > gpurun_out/pytest_ae.log 2>&1; echo rc=$?; tail -5 gpurun_out/pytest_ae.log
echo "== prof ae bf16"; timeout 300 python tools/prof_ae.py --rows 1010000 --precision bf16 --iters 6 2>&1 | tail -3
echo "== no prefetch"; LK_AE_PF=0 timeout 300 python tools/prof_ae.py --rows 1010000 --precision bf16 --iters 5 2>&1 | tail -2
echo "== 2 stages"; LK_AE_STAGES=2 timeout 300 python tools/prof_ae.py --rows 1010000 --precision bf16 --iters 5 2>&1 | tail -2
CMD="python tools/prof_ae.py --rows 1010000 --precision bf16 --iters 2"
ncu --set full --clock-control none --import-source on -k regex:ae_pair -s 1 -c 1 -f -o gpurun_out/prof_ae_pair3 $CMD > gpurun_out/prof_ae_ncu.log 2>&1; echo "ncu rc=$?"
echo "== c1 selector experiment"
python tools/prof_case.py --rows 20000 --batch 10000 --iters 4 2>&1 | tail -2
LK_KSEL_BUF=1 python tools/prof_case.py --rows 20000 --batch 10000 --iters 4 2>&1 | tail -2
echo "== d64 selector experiment"
python tools/prof_case.py --rows 1000000 --dim 64 --batch 10000 --iters 3 2>&1 | tail -1
LK_KSEL_BUF=1 python tools/prof_case.py --rows 1000000 --dim 64 --batch 10000 --iters 3 2>&1 | tail -1
echo "== c1 launch list"
CMD2="python tools/prof_case.py --rows 20000 --batch 10000 --iters 3"
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/c1_launches.csv $CMD2 > /dev/null 2>&1; python tools/summarize_launches.py gpurun_out/c1_launches.csv 2>/dev/null | head -12
