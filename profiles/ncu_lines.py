#!/usr/bin/env python
"""Per CUDA source line: share of executed warp instructions and of stall samples (needs -lineinfo + --import-source).
usage: python profiles/ncu_lines.py report.ncu-rep [top_n] [divisor]"""
import csv
import subprocess
import sys

rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 30
div = float(sys.argv[3]) if len(sys.argv) > 3 else None
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv"], capture_output=True, text=True).stdout
fname, hdr, L = "", None, []
for r in csv.reader(out.splitlines()):
    if len(r) == 2 and r[0] == "File Name":
        fname = r[1].split("/")[-1]
        continue
    if len(r) > 6 and r[0] == "Line No":
        hdr = r
        ie, st, te = hdr.index("Instructions Executed"), hdr.index("Warp Stall Sampling (All Samples)"), hdr.index("Thread Instructions Executed")
        continue
    if hdr and len(r) > 6 and r[0] not in ("", "Line No"):
        try:
            L.append((int(r[ie]), int(r[st]), int(r[te]), f"{fname}:{r[0]}", r[1].strip()))
        except ValueError:
            pass
tot, ts = sum(x[0] for x in L), sum(x[1] for x in L)
print(f"{tot} warp instructions, {ts} stall samples" + (f", {tot / div:.0f} per unit" if div else ""))
for n, s, t, loc, src in sorted(L, key=lambda x: -(x[0] / max(tot, 1) + x[1] / max(ts, 1)))[:top]:
    print(f"{100 * n / tot:5.1f}% inst {100 * s / max(ts, 1):5.1f}% stall  lanes {t / max(n, 1):4.1f}  {loc:22s} {src[:110]}")
