#!/bin/bash
# Build liblatentknn.so with extra -D flags on lk_search_umma.cu only and park it under tools/_bin/ (A/B aid):
#   tools/build_variant.sh <name> [-DKNOB=V ...]      -> tools/_bin/lib_<name>.so ; the in-tree build is restored by
#   touching the source afterwards (run `python -m latent_rag_b200.build` to get the default back)
set -eu
name=$1; shift
mkdir -p tools/_bin
touch latent_rag_b200/csrc/lk_search_umma.cu
LK_NVCC_EXTRA="$*" python -m latent_rag_b200.build > /dev/null
cp latent_rag_b200/liblatentknn.so tools/_bin/lib_${name}.so
touch latent_rag_b200/csrc/lk_search_umma.cu
