#!/bin/bash
cd "$(dirname "$0")/.."
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -p no:cacheprovider -k "search" > gpurun_out/r2_pytest_search.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2_pytest_search.log
timeout 300 python tools/search_ab.py 16384 > gpurun_out/r2_search_ab.json 2> gpurun_out/r2_search_ab.err; echo "ab rc=$?"
VNLB_SEARCH_PATH=2 timeout 600 ncu --set full --clock-control none --import-source on -k regex:search_ -c 4 -o gpurun_out/prof_r2_quad -f python tools/run_kernels.py search 2048 > gpurun_out/r2_ncu_quad.log 2>&1; echo "ncu quad rc=$?"
