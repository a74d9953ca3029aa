#!/usr/bin/env python
"""Per-source-line executed warp-instruction counts of one kernel from an .ncu-rep (captured with --import-source on).
usage: tools/ncu_lines.py REPORT.ncu-rep KERNEL_REGEX [min_share_percent]"""
import csv, subprocess, sys
rep, kern = sys.argv[1], sys.argv[2]
minp = float(sys.argv[3]) if len(sys.argv) > 3 else 0.5
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + kern, "--print-source", "cuda,sass"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = None
lines = []
fname = ""
for r in rows:
    if len(r) >= 2 and r[0] == "File Path":
        fname = r[1].split("/")[-1]
    if len(r) > 8 and r[0] == "Line No":
        hdr = r
        ie = hdr.index("Instructions Executed"); te = hdr.index("Thread Instructions Executed"); ss = hdr.index("# Samples")
        continue
    if hdr and len(r) > ie and r[0].strip().isdigit():
        try:
            lines.append((fname, int(r[0]), r[1].strip(), int(r[ie]), int(r[te]), int(r[ss])))
        except ValueError:
            pass
tot = sum(l[3] for l in lines); tots = sum(l[5] for l in lines)
print(f"# {kern}: {tot} warp-instructions, {tots} samples")
for f, n, src, i, t, s in lines:
    if 100.0 * i / tot >= minp or 100.0 * s / max(tots, 1) >= minp:
        print(f"{f}:{n:4d} {100.0*i/tot:5.1f}% inst {100.0*s/max(tots,1):5.1f}% smp  thr/inst {t/max(i,1):4.1f}  {src[:110]}")
# totals per file
from collections import defaultdict
pf = defaultdict(int)
for f, n, src, i, t, s in lines:
    pf[f] += i
print("# per file:", {k: f"{100.0*v/tot:.1f}%" for k, v in pf.items()})
if len(sys.argv) > 4:  # buckets "a-b,c-d,..." of the main file
    main = max(pf, key=pf.get)
    for rng in sys.argv[4].split(","):
        a, b = map(int, rng.split("-"))
        v = sum(l[3] for l in lines if l[0] == main and a <= l[1] <= b); tt = sum(l[4] for l in lines if l[0] == main and a <= l[1] <= b)
        sm = sum(l[5] for l in lines if l[0] == main and a <= l[1] <= b)
        print(f"  {main}:{a}-{b}: {100.0*v/tot:.1f}% inst, {100.0*sm/max(tots,1):.1f}% samples, thr/inst {tt/max(v,1):.1f}")
