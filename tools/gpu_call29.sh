#!/bin/bash
cd "$(dirname "$0")/.."
python tools/run_kernels.py fused 2048 > /dev/null 2>&1 || exit 1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"bayes_kernel" -c 3 -o /tmp/prof_lsu -f python tools/run_kernels.py fused 2048 > gpurun_out/r2_ncu_lsu.log 2>&1; echo "ncu rc=$?"
for skip in 0 2; do
  echo "=== bayes_kernel launch $skip" >> gpurun_out/r2_lsu_lines_bayes.txt
  python tools/ncu_lsu_lines.py /tmp/prof_lsu.ncu-rep "bayes_kernel" $skip >> gpurun_out/r2_lsu_lines_bayes.txt 2>&1
done
