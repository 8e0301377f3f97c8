#!/bin/bash
# round-2 GPU call: search kernel variants -- parity tests, then the A/B timing
cd "$(dirname "$0")/.."
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -p no:cacheprovider -k "search" > gpurun_out/r2_pytest_search.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/r2_pytest_search.log
timeout 300 python tools/search_ab.py 16384 > gpurun_out/r2_search_ab.json 2> gpurun_out/r2_search_ab.err; echo "ab rc=$?"; cat gpurun_out/r2_search_ab.json; tail -5 gpurun_out/r2_search_ab.err
