#!/bin/bash
# every switchable kernel path through the parity tests
cd "$(dirname "$0")/.."
for env in "VNLB_COV4=0" "VNLB_SEARCH_PATH=1" "VNLB_TAIL2=0" "VNLB_BAYES_SPLIT=0"; do
  env $env timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -p no:cacheprovider -k "bayes or fused or e2e or search or covariance or graph" > gpurun_out/r2_pytest_env.log 2>&1; echo "$env rc=$? $(tail -1 gpurun_out/r2_pytest_env.log)"
done
