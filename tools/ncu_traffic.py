"""Reads an `ncu --set full` report of `tools/run_kernels.py fused <rows>` (one round of <rows> groups per VNLB step) and
writes the summary bench.py's roofline reads (profiles/ncu_traffic.json): DRAM bytes per group of the fused Bayes stage
(dram__bytes_read.sum + dram__bytes_write.sum over its kernels, divided by the groups of the captured launch) and the
pipe utilisation of every kernel.   usage: python tools/ncu_traffic.py REPORT.ncu-rep ROWS [out.json]"""
import csv
import io
import json
import os
import re
import subprocess
import sys

rep, rows = sys.argv[1], int(sys.argv[2])
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
out_path = sys.argv[3] if len(sys.argv) > 3 else os.path.join(ROOT, "profiles", "ncu_traffic.json")
txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
r = list(csv.reader(io.StringIO(txt)))
h, units, data = r[0], r[1], r[2:]


def val(row, key):
    if key not in h:
        return None
    i = h.index(key)
    try:
        v = float(row[i].replace(",", ""))
    except ValueError:
        return None
    u = units[i].lower()
    scale = {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9, "us": 1e-6, "ms": 1e-3, "ns": 1e-9, "s": 1}.get(u, 1)
    return v * scale


kernels = []
for row in data:
    name = re.sub(r"\(.*", "", row[h.index("Kernel Name")]).replace("void ", "").replace("vnlb::", "")
    kernels.append(dict(
        name=name, grid=val(row, "launch__grid_size"), block=val(row, "launch__block_size"),
        seconds=val(row, "gpu__time_duration.sum"), registers=val(row, "launch__registers_per_thread"),
        dram_read=val(row, "dram__bytes_read.sum"), dram_write=val(row, "dram__bytes_write.sum"),
        fma_pipe_pct=val(row, "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active"),
        fma_inst_pct=val(row, "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active"),
        issue_pct=val(row, "smsp__issue_active.avg.pct_of_peak_sustained_active"),
        lsu_wavefront_pct=val(row, "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed"),
        tensor_pipe_pct=val(row, "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"),
        dram_pct=val(row, "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"),
        warps_active_pct=val(row, "sm__warps_active.avg.pct_of_peak_sustained_active")))
# the capture holds step 1 (search, cov_tridiag, tail, tail, bayes<direct>) then step 2 (search, gram_tridiag, tail, bayes<gram>)
def stage(pred):
    ks = [k for k in kernels if pred(k["name"])]
    return ks
s1, s2, seen_gram = [], [], False
for k in kernels:
    if "search" in k["name"]:
        continue
    if "gram_tridiag" in k["name"]:
        seen_gram = True
    (s2 if seen_gram else s1).append(k)
def summarise(ks):
    calls = max(1, sum(1 for k in ks if "cov_tridiag" in k["name"] or "gram_tridiag" in k["name"]))   # C-ABI calls captured
    byts = sum((k["dram_read"] or 0) + (k["dram_write"] or 0) for k in ks) / calls
    t = sum(k["seconds"] or 0 for k in ks) / calls
    w = lambda key: sum((k[key] or 0) * (k["seconds"] or 0) for k in ks) / max(t * calls, 1e-30)
    return byts / rows, dict(kernel_seconds=t, fma_pipe_pct=w("fma_pipe_pct"), issue_pct=w("issue_pct"),
                             lsu_wavefront_pct=w("lsu_wavefront_pct"), tensor_pipe_pct=w("tensor_pipe_pct"), dram_pct=w("dram_pct"),
                             note="time-weighted over the stage's kernels")
b1, p1 = summarise(s1)
b2, p2 = summarise(s2)
rec = dict(source="ncu --set full --clock-control none on `python tools/run_kernels.py fused %d` (%s), dram__bytes_read.sum + "
                  "dram__bytes_write.sum of the stage's kernels / %d groups" % (rows, os.path.basename(rep), rows),
           rows=rows, bayes_step1_dram_bytes_per_group=b1, bayes_step2_dram_bytes_per_group=b2,
           bayes_step1_pipes=p1, bayes_step2_pipes=p2, kernels=kernels)
json.dump(rec, open(out_path, "w"), indent=1)
print(json.dumps({k: v for k, v in rec.items() if k != "kernels"}, indent=1))
