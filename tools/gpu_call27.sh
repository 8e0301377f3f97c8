#!/bin/bash
cd "$(dirname "$0")/.."
python tools/run_kernels.py fused 2048 > /dev/null 2>&1 || exit 1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"cov_tridiag|bayes_kernel|gram_tridiag" -c 6 -o /tmp/prof_lsu -f python tools/run_kernels.py fused 2048 > gpurun_out/r2_ncu_lsu.log 2>&1; echo "ncu rc=$?"
for k in cov_tridiag4 "bayes_kernel<\(bool\)1, \(bool\)0" "bayes_kernel<\(bool\)1, \(bool\)1" gram_tridiag; do
  echo "=== $k" >> gpurun_out/r2_lsu_lines.txt
  python tools/ncu_lsu_lines.py /tmp/prof_lsu.ncu-rep "$k" 0 >> gpurun_out/r2_lsu_lines.txt 2>&1
done
python tools/ncu_summary.py /tmp/prof_lsu.ncu-rep | cut -c1-220 | sed -n 1,6p
