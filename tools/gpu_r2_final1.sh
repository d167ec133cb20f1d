#!/bin/bash
# Round-2 evidence visit (one GPU): full GPU test suite, smoke, bench (all configs) + reference arm, ncu captures.
set -u
mkdir -p gpurun_out
echo "== pytest all"; timeout 1800 python -m pytest tests -m gpu -q > gpurun_out/pytest_all.log 2>&1; echo rc=$?; tail -4 gpurun_out/pytest_all.log | cut -c1-200
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo rc=$?; tail -2 gpurun_out/smoke.log
echo "== reference arm"; timeout 900 python bench.py --impl reference --steps 5 --warmup 2 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo rc=$?
echo "== bench"; timeout 1500 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_full.json 2> gpurun_out/bench_full.err; echo rc=$?; cut -c1-200 gpurun_out/bench_full.json; tail -3 gpurun_out/bench_full.err
echo "== ncu ae"
CMD="python tools/prof_ae.py --rows 1010000 --precision bf16 --iters 2"
$CMD > gpurun_out/prof_ae_plain.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:ae_pair -s 1 -c 1 -f -o gpurun_out/prof_ae_pair_final $CMD > gpurun_out/prof_ae_ncu.log 2>&1; echo "rc=$?"; tail -1 gpurun_out/prof_ae_plain.log
echo "== ncu d64"
CMD="python tools/prof_case.py --rows 1000000 --dim 64 --batch 10000 --iters 2"
$CMD > gpurun_out/prof_d64_plain.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:umma_search -s 1 -c 1 -f -o gpurun_out/prof_d64_r02 $CMD > gpurun_out/prof_d64_ncu.log 2>&1; echo "rc=$?"
echo "== ncu fp32 planes (c1 fp32)"
cat > /tmp/prof_fp32.py <<'PY'
import sys, torch; sys.path.insert(0, '.')
import latent_rag_b200 as lrb
g = torch.Generator(device="cuda").manual_seed(1)
ix = lrb.ExactIndex(384, 20000, metric="cosine", storage="fp32"); ix.add(torch.randn((20000, 384), generator=g, device="cuda"))
q = torch.randn((10000, 384), generator=g, device="cuda"); ix.set_timing(True)
for it in range(3):
    ix.search(q, 10, device_out=True); print("iter", it, ix.last_timing(), flush=True)
PY
python /tmp/prof_fp32.py > gpurun_out/prof_fp32_plain.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:umma_search -s 1 -c 1 -f -o gpurun_out/prof_fp32_r02 python /tmp/prof_fp32.py > gpurun_out/prof_fp32_ncu.log 2>&1; echo "rc=$?"; tail -1 gpurun_out/prof_fp32_plain.log
echo "== launch list of a bench run (10M rows, every config)"
BENCH="python bench.py --rows 10000000 --steps 2 --warmup 3"
$BENCH > gpurun_out/bench_for_launches.json 2> gpurun_out/bench_for_launches.err && ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/launches_r02.csv $BENCH > gpurun_out/launches_run.log 2>&1; echo "rc=$?"
