#!/bin/bash
# Wiener filter on the tensor cores (VNLB_FILTER_MMA=1) against the FFMA2 filter: parity tests, then the Bayes microbenchmark
cd "$(dirname "$0")/.."
VNLB_FILTER_MMA=1 timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -p no:cacheprovider -k "bayes or fused or tensor" > gpurun_out/r2_pytest_mma.log 2>&1; echo "pytest(mma) rc=$?"; tail -1 gpurun_out/r2_pytest_mma.log
for v in 0 1; do
VNLB_FILTER_MMA=$v python tools/microbench.py 4096 16384 2>/dev/null | python -c "
import json,sys; d=json.load(sys.stdin); print('mma=$v', {k: round(v['ms'],3) for k,v in d['bayes_config4'].items()})"
done
