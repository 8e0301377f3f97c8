#!/bin/bash
cd "$(dirname "$0")/.."
VNLB_COV4=1 timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -p no:cacheprovider -k "bayes or fused or covariance or e2e" > gpurun_out/r2_pytest_bayes.log 2>&1; echo "pytest(cov4) rc=$?"; tail -2 gpurun_out/r2_pytest_bayes.log
VNLB_COV4=0 python tools/microbench.py 4096 16384 > gpurun_out/r2_mb_cov0.json 2>&1
VNLB_COV4=1 python tools/microbench.py 4096 16384 > gpurun_out/r2_mb_cov1.json 2>&1
python - <<'PY'
import json
for f in ("gpurun_out/r2_mb_cov0.json", "gpurun_out/r2_mb_cov1.json"):
    try:
        d = json.load(open(f)); print(f, {k: round(v["ms"], 3) for k, v in d["bayes_config4"].items()})
    except Exception as e:
        print(f, "unreadable", e)
PY
