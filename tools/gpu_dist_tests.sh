#!/bin/bash
# 2 GPUs: NCCL tests of the band + halo path
cd "$(dirname "$0")/.."
timeout 600 python -m pytest tests/test_dist_gpu.py -m gpu -q -p no:cacheprovider > gpurun_out/r2_pytest_dist.log 2>&1; echo "pytest dist rc=$?"; tail -3 gpurun_out/r2_pytest_dist.log
