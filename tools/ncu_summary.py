"""Markdown table of key counters per launch of an ncu report: python tools/ncu_summary.py REPORT [kernel-regex]"""
import csv, io, subprocess, sys, re
rep = sys.argv[1]
kre = re.compile(sys.argv[2]) if len(sys.argv) > 2 else None
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
r = list(csv.reader(io.StringIO(out)))
h, units, rows = r[0], r[1], r[2:]
keys = ["launch__grid_size", "launch__block_size", "gpu__time_duration.sum", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"]
sel = [v for v in rows if kre is None or kre.search(v[h.index("Kernel Name")])]
names = [re.sub(r"\(.*", "", v[h.index("Kernel Name")]).replace("void ", "").replace("vnlb::", "") for v in sel]
print("| metric | " + " | ".join("%d: %s" % (i, n) for i, n in enumerate(names)) + " |")
print("|---|" + "---|" * len(sel))
for k in keys:
    if k not in h:
        continue
    i = h.index(k)
    def fmt(x):
        try:
            f = float(x.replace(",", ""))
            return "%.4g" % f
        except ValueError:
            return x
    print("| %s (%s) | " % (k, units[i]) + " | ".join(fmt(v[i]) for v in sel) + " |")
