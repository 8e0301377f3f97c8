#!/bin/bash
# round-2 profile evidence: per-kernel counters of one round (2048 groups per step), launch list of a short bench call.
# The .ncu-rep is summarised ON the box (gpurun_out/ is capped at 64 MiB) and only kept when small.
cd "$(dirname "$0")/.."
python tools/run_kernels.py fused 2048 > gpurun_out/r2_plain_kernels.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"search_|cov_tridiag|tridiag_tail|gram_tridiag|bayes_kernel" -c 16 -o /tmp/prof_r2b -f python tools/run_kernels.py fused 2048 > gpurun_out/r2_ncu_full.log 2>&1; echo "ncu full rc=$?"
python tools/ncu_summary.py /tmp/prof_r2b.ncu-rep > gpurun_out/r2_kernel_table.md 2> gpurun_out/r2_kernel_table.err
python tools/ncu_traffic.py /tmp/prof_r2b.ncu-rep 2048 gpurun_out/r2_ncu_traffic.json > gpurun_out/r2_ncu_traffic.log 2>&1
ncu -i /tmp/prof_r2b.ncu-rep --page raw --csv > gpurun_out/r2_raw.csv 2>/dev/null
for k in cov_tridiag bayes_kernel gram_tridiag search_quad; do
  for skip in 0 1; do python tools/ncu_phase.py /tmp/prof_r2b.ncu-rep $k vnlb_b200/csrc/$( [ $k = search_quad ] && echo search.cu || echo bayes_tridiag.cu ) $skip > gpurun_out/r2_phase_${k}_$skip.txt 2>&1; done
done
ls -la /tmp/prof_r2b.ncu-rep
sz=$(stat -c %s /tmp/prof_r2b.ncu-rep); if [ "$sz" -lt 40000000 ]; then cp /tmp/prof_r2b.ncu-rep gpurun_out/; fi
python bench.py --quick --frames 8 --steps 1 --warmup 1 > gpurun_out/r2_plain_bench.log 2>&1 && \
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/r2_launches_bench.csv python bench.py --quick --frames 8 --steps 1 --warmup 1 > gpurun_out/r2_ncu_launch.log 2>&1; echo "ncu launches rc=$?"
du -sh gpurun_out
