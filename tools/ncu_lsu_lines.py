"""Top source lines of a kernel by shared-memory wavefronts (and the excess over the conflict-free ideal):
python tools/ncu_lsu_lines.py REPORT KERNEL_REGEX [launch_skip]"""
import csv, io, subprocess, sys
rep, kre = sys.argv[1], sys.argv[2]
skip = sys.argv[3] if len(sys.argv) > 3 else "0"
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "--kernel-name", "regex:" + kre,
                      "--launch-skip", skip, "--launch-count", "1"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Line No")
hdr = rows[hi]
iW, iE, iI, iS = hdr.index("L1 Wavefronts Shared"), hdr.index("L1 Wavefronts Shared Excessive"), hdr.index("Instructions Executed"), hdr.index("# Samples")
lines = []
for r in rows[hi + 1:]:
    if len(r) > iW and r[0] != "":
        try:
            lines.append((int(r[0]), r[1][:100], float(r[iW] or 0), float(r[iE] or 0), float(r[iI] or 0), float(r[iS] or 0)))
        except ValueError:
            pass
tw, te, ti = sum(l[2] for l in lines), sum(l[3] for l in lines), sum(l[4] for l in lines)
print("shared wavefronts %.3e (excessive %.3e = %.1f%%), warp instr %.3e" % (tw, te, 100 * te / max(tw, 1), ti))
for l in sorted(lines, key=lambda x: -x[2])[:28]:
    print("%5d %5.1f%% wf (%4.1f%% excess) %5.1f%% ins %5.1f%% smp  %s" % (l[0], 100 * l[2] / tw, 100 * l[3] / max(l[2], 1), 100 * l[4] / ti, 100 * l[5] / max(sum(x[5] for x in lines), 1), l[1]))
