#!/bin/bash
# epilogue phase counters (LK_EPI_PROF builds parked by tools/build_variant.sh) for two variants
set -u
mkdir -p gpurun_out
for v in ${LK_AB_VARIANTS:-base_prof share_prof}; do
  echo "#### $v"
  for c in "--rows 20000 --batch 10000" "--rows 1000000 --dim 64 --batch 10000" "--rows 4000000 --batch 4096"; do
    echo "case $c"
    LK_UMMA_DUMP=/tmp/epi.bin python tools/ab_old_lib.py tools/_bin/lib_$v.so $c --iters 2 2>&1 | tail -1
    python tools/epi_prof.py /tmp/epi.bin
  done
done
