#!/bin/bash
cd "$(dirname "$0")/.."
for env in "VNLB_BAYES_SPLIT=0" "VNLB_BAYES_SPLIT=1"; do
  env $env timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -p no:cacheprovider -k "bayes or fused or e2e or covariance or graph or flat" > gpurun_out/r2_pytest_env.log 2>&1; echo "$env rc=$? $(tail -1 gpurun_out/r2_pytest_env.log)"
done
python tools/microbench.py 4096 16384 2>/dev/null | python -c "
import json,sys; d=json.load(sys.stdin); print({k: round(v['ms'],3) for k,v in d['bayes_config4'].items()})"
