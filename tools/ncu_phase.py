"""Per-phase breakdown of an ncu report for a kernel: python tools/ncu_phase.py REPORT KERNEL_REGEX SOURCE_FILE [launch_index]"""
import csv, subprocess, sys, io, re
rep, kre, srcfile = sys.argv[1], sys.argv[2], sys.argv[3]
skip = sys.argv[4] if len(sys.argv) > 4 else "0"
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "--kernel-name", "regex:" + kre,
                      "--launch-skip", skip, "--launch-count", "1"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Line No")
hdr = rows[hi]
iS, iI = hdr.index("# Samples"), hdr.index("Instructions Executed")
lines = []
for r in rows[hi + 1:]:
    if len(r) > iI and r[0] != "":
        try: lines.append((int(r[0]), r[1][:110], float(r[iS] or 0), float(r[iI] or 0)))
        except ValueError: pass
tot_s = sum(l[2] for l in lines); tot_i = sum(l[3] for l in lines)
src = open(srcfile).read().split("\n")
marks = [(i + 1, l.strip(" /-")) for i, l in enumerate(src) if re.search(r"// -{10,} ", l)]
bounds = [(1, "prologue/helpers")] + marks
print("total samples %d, warp instr %.3e" % (tot_s, tot_i))
for bi, (b, nm) in enumerate(bounds):
    e = bounds[bi + 1][0] - 1 if bi + 1 < len(bounds) else 10 ** 9
    s = sum(l[2] for l in lines if b <= l[0] <= e); i = sum(l[3] for l in lines if b <= l[0] <= e)
    print("%-60s samples %5.1f%%  instr %5.1f%%  %8.3e" % (nm[:60], 100 * s / tot_s, 100 * i / tot_i, i))
print("--- top lines by samples")
for l in sorted(lines, key=lambda x: -x[2])[:22]:
    print("%4d %5.1f%% smp %5.1f%% ins  %s" % (l[0], 100 * l[2] / tot_s, 100 * l[3] / tot_i, l[1]))
