#!/bin/bash
# Build bottleneck-analysis variants of the library into rcdms_b200/_Cx<N>/ (git-ignored): RCDM_GEMM_EXPERIMENT=N
for n in "$@"; do
  RCDM_BUILD_DIR=$PWD/rcdms_b200/_Cx$n RCDM_EXTRA_NVCC_FLAGS="-DRCDM_GEMM_EXPERIMENT=$n" python -m rcdms_b200.build --force
done
