#!/bin/bash
# TMA tile-load probe (profiles/r2_search_summary.md): build here (nvcc cross-compiles), run on the GPU box:
#   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o tools/experimental/tma_probe tools/experimental/tma_probe.cu
#   gpurun -- 'bash tools/gpu_tma_probe.sh'
cd "$(dirname "$0")/.."
[ -x tools/experimental/tma_probe ] || nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o tools/experimental/tma_probe tools/experimental/tma_probe.cu
for v in 0 1 2 3 4 5; do timeout 60 ./tools/experimental/tma_probe $v 2>&1 | tr '\n' ' '; echo; done
