#!/usr/bin/env python
"""Per-kernel SASS mnemonic counts of libbioranges.so (cuobjdump -sass | c++filt): which kernels use which hardware path.
usage: python profiles/sass_summary.py [lib.so]  > profiles/r2_sass_summary.txt"""
import collections, re, subprocess, sys
lib = sys.argv[1] if len(sys.argv) > 1 else "datafusion-bio-functions_b200/libbioranges.so"
sass = subprocess.run(f"cuobjdump -sass {lib} | c++filt", shell=True, capture_output=True, text=True).stdout
cols = [("UBLKCP", r"\bUBLKCP"), ("ARRIVE", r"SYNCS\.ARRIVE"), ("PHCHK", r"SYNCS\.PHASECHK"), ("UTMALDG", r"UTMALDG"), ("UTCMMA", r"UTC\w*MMA"),
        ("MATCH", r"\bMATCH"), ("ATOMS", r"\bATOMS"), ("ATOMG", r"\bATOMG|\bRED\."), ("LDG128", r"LDG\.E\.(\w+\.)*128"), ("LDG256", r"LDG\.E\.(\w+\.)*256"),
        ("STG128", r"STG\.E\.(\w+\.)*128")]
kern, counts = None, collections.OrderedDict()
for line in sass.splitlines():
    m = re.match(r"\s*Function : (.*)", line)
    if m:
        name = re.sub(r"\(.*", "", m.group(1)).replace("brg::", "").replace("void ", "")
        kern = name
        counts.setdefault(kern, collections.Counter())
        continue
    if kern and re.match(r"\s*/\*[0-9a-f]{4,6}\*/", line):
        c = counts[kern]
        c["instr"] += 1
        for n, pat in cols:
            if re.search(pat, line):
                c[n] += 1
print("SASS of", lib, "(cuobjdump -sass, sm_100a), per kernel: instruction count and the mnemonics that show which hardware path a kernel uses.")
print("UBLKCP = 1-D TMA bulk copy (cp.async.bulk), SYNCS.ARRIVE.TRANS / SYNCS.PHASECHK = mbarrier arrive-with-transaction-count / parity wait,")
print("ATOMG = global atomics / reductions (tickets, cursors), LDG256 = 256-bit loads.  No UTMALDG (tensor-map TMA: every stream here is 1-D) and no")
print("UTC*MMA (tcgen05: nothing on this path is a contraction).\n")
print(f"{'kernel':58s}" + f"{'instr':>7s}" + "".join(f"{n:>8s}" for n, _ in cols))
for k in sorted(counts):
    c = counts[k]
    print(f"{k[:58]:58s}{c['instr']:7d}" + "".join(f"{c[n]:8d}" for n, _ in cols))
tot = collections.Counter()
for c in counts.values():
    tot.update(c)
print(f"\ntotals: kernels {len(counts)}, instructions {tot['instr']}, " + ", ".join(f"{n} {tot[n]}" for n, _ in cols[:5]))
print("""
Where the TMA path is and is not used, and why:
 * round 1's chunk-wise probe kernels (count / coverage / nearest_k1 _kernel_tiled, pairs_ranges_kernel_tiled) stream the PROBE columns
   through per-warp shared-memory rings filled by UBLKCP and completed on mbarriers: long contiguous columns read exactly once;
 * round 2's windowed kernels (pairs_join_kernel, nearest_k1_kernel_win, rank2_kernel_win) copy a chunk's WINDOW of the sorted build
   arrays (and, for nearest / coverage, of the 16-byte records) into shared memory by UBLKCP on two mbarriers, so that the second
   copy lands while the first is searched; their probe rows come through plain coalesced loads (32 consecutive rows per instruction);
 * K1 (radix_*), K2 (scans, merge, directories, count bins) and the per-row kernels have no UBLKCP: their global accesses are either
   register-resident tiles loaded with plain coalesced loads right before use (sort: 16 rows per thread, all loads issued up front),
   data-dependent gathers (one record per row; LDG.E.256 for the 32-byte count-bin records) or scattered stores (sort) -- none of
   them has a producer/consumer stage a bulk copy could decouple.
 * global atomics: the chunk ticket and the output cursor of pairs_join_kernel, the onesweep tile tickets of radix_scatter_kernel,
   the to-do lists of the chunk-wise kernels.""")
