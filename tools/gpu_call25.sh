#!/bin/bash
cd "$(dirname "$0")/.."
VNLB_COV4=1 python tools/run_kernels.py fused 2048 > /dev/null 2>&1 || exit 1
VNLB_COV4=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:"cov_tridiag" -c 1 -o /tmp/prof_cov4 -f python tools/run_kernels.py fused 2048 > gpurun_out/r2_ncu_cov4.log 2>&1; echo "ncu rc=$?"
python tools/ncu_summary.py /tmp/prof_cov4.ncu-rep > gpurun_out/r2_cov4_table.md 2>&1
python tools/ncu_phase.py /tmp/prof_cov4.ncu-rep cov_tridiag4 vnlb_b200/csrc/bayes_tridiag.cu 0 > gpurun_out/r2_cov4_phase.txt 2>&1
head -24 gpurun_out/r2_cov4_phase.txt
