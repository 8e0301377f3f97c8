#!/bin/bash
# Builds vnlb_b200/libvnlb_b200.so (sm_100a) and the CPU oracle.
set -e
cd "$(dirname "$0")"
SRC="vnlb_b200/csrc/pixel_ops.cu vnlb_b200/csrc/search.cu vnlb_b200/csrc/bayes.cu vnlb_b200/csrc/bayes_jacobi.cu vnlb_b200/csrc/bayes_tridiag.cu vnlb_b200/csrc/aggregate.cu"
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 \
     -Xcompiler -fPIC -shared ${NVCC_EXTRA} -o vnlb_b200/libvnlb_b200.so $SRC
make -s -C oracle
