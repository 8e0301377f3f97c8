#!/bin/bash
cd "$(dirname "$0")/.."
VNLB_SEARCH_PATH=2 python tools/run_kernels.py search 2048 > gpurun_out/r2_plain_search.log 2>&1 || exit 1
VNLB_SEARCH_PATH=2 timeout 600 ncu --set full --clock-control none --import-source on -k regex:search_ -c 4 -o gpurun_out/prof_r2_quad -f python tools/run_kernels.py search 2048 > gpurun_out/r2_ncu_quad.log 2>&1; echo "ncu quad rc=$?"
VNLB_SEARCH_PATH=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:search_ -c 4 -o gpurun_out/prof_r2_tiled -f python tools/run_kernels.py search 2048 > gpurun_out/r2_ncu_tiled.log 2>&1; echo "ncu tiled rc=$?"
ls -la gpurun_out/*.ncu-rep
