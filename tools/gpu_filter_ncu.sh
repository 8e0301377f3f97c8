#!/bin/bash
cd "$(dirname "$0")/.."
VNLB_FILTER_MMA=1 python tools/run_kernels.py fused 2048 > /dev/null 2>&1 || exit 1
VNLB_FILTER_MMA=1 timeout 900 ncu --set full --clock-control none -k regex:"bayes_kernel" -c 3 -o /tmp/prof_mma -f python tools/run_kernels.py fused 2048 > gpurun_out/r2_ncu_mma.log 2>&1; echo "ncu rc=$?"
python tools/ncu_summary.py /tmp/prof_mma.ncu-rep > gpurun_out/r2_filter_mma_table.md 2>&1
cut -c1-140 gpurun_out/r2_filter_mma_table.md | sed -n 1,22p
