#!/usr/bin/env python
"""Summarise an .ncu-rep (read here, no GPU needed): headline metrics, stall mix, opcode mix, hottest SASS lines.
usage: python profiles/ncu_summary.py gpurun_out/prof.ncu-rep [warp_iterations_divisor]"""
import collections
import csv
import subprocess
import sys

rep = sys.argv[1]
div = float(sys.argv[2]) if len(sys.argv) > 2 else None
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units = rows[0], rows[1]
KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block_dynamic",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum"]
for r in rows[2:]:
    print("kernel:", r[hdr.index("Kernel Name")][:90])
    for k in KEYS:
        if k in hdr:
            print(f"  {k:75s} {r[hdr.index(k)]} {units[hdr.index(k)]}")
    st = [(float(r[i]), h) for i, h in enumerate(hdr) if h.startswith("smsp__average_warps_issue_stalled") and h.endswith("_per_issue_active.ratio") and "not_issued" not in h]
    print("  stall mix (warps per issue):", ", ".join(f"{h.split('stalled_')[1].split('_per_')[0]}={v:.2f}" for v, h in sorted(st, reverse=True)[:7]))
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(src.splitlines()))
for i, r in enumerate(rows):
    if "Source" in r and "Instructions Executed" in r:
        h = i
        break
hdr = rows[h]
si, ie, st = hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("Warp Stall Sampling (All Samples)")
tot = tst = 0
ops = collections.Counter()
lines = []
for r in rows[h + 1:]:
    try:
        n, s = int(r[ie]), int(r[st])
    except (ValueError, IndexError):
        continue
    tot += n
    tst += s
    t = r[si].split()
    op = t[1] if t[0].startswith("@") else t[0]
    ops[op.split(".")[0]] += n
    lines.append((n, s, r[si]))
print(f"SASS: {tot} warp instructions, {tst} stall samples" + (f", {tot / div:.1f} per unit" if div else ""))
print("  opcode mix:", ", ".join(f"{k} {100 * v / tot:.1f}%" for k, v in ops.most_common(10)))
print("  hottest lines by stall samples:")
for n, s, t in sorted(lines, key=lambda x: -x[1])[:16]:
    print(f"    {100 * s / max(tst, 1):5.1f}%  x{(n / div if div else n):9.2f}  {t[:100]}")
