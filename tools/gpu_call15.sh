#!/bin/bash
cd "$(dirname "$0")/.."
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -p no:cacheprovider -k "bayes or fused or e2e or covariance" > gpurun_out/r2_pytest_bayes.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2_pytest_bayes.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"cov_tridiag" -c 2 -o /tmp/prof_cov4 -f python tools/run_kernels.py fused 2048 > gpurun_out/r2_ncu_cov4.log 2>&1; echo "ncu rc=$?"
python tools/ncu_summary.py /tmp/prof_cov4.ncu-rep > gpurun_out/r2_cov4_table.md 2>&1
python tools/ncu_phase.py /tmp/prof_cov4.ncu-rep cov_tridiag4 vnlb_b200/csrc/bayes_tridiag.cu 0 > gpurun_out/r2_cov4_phase.txt 2>&1
ncu -i /tmp/prof_cov4.ncu-rep --page raw --csv > gpurun_out/r2_cov4_raw.csv 2>/dev/null
