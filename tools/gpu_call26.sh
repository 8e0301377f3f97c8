#!/bin/bash
cd "$(dirname "$0")/.."
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -p no:cacheprovider -k "bayes or fused or covariance or e2e" > gpurun_out/r2_pytest_bayes.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/r2_pytest_bayes.log
python tools/microbench.py 4096 16384 > gpurun_out/r2_mb_tw.json 2>&1
python - <<'PY'
import json
d = json.load(open("gpurun_out/r2_mb_tw.json")); print({k: round(v["ms"], 3) for k, v in d["bayes_config4"].items()})
PY
