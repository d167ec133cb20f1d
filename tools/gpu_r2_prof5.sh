#!/bin/bash
# epilogue phase counters (LK_EPI_PROF build parked by tools/build_variant.sh prof -DLK_EPI_PROF) across batch sizes
set -u
for b in 256 1024 4096; do
  echo "case 20M x 384, $b queries"
  LK_UMMA_DUMP=/tmp/epi.bin python tools/ab_old_lib.py tools/_bin/lib_prof.so --rows 20000000 --batch $b --iters 2 2>&1 | tail -1
  python tools/epi_prof.py /tmp/epi.bin
done
