"""count_overlaps kernel time on the headline table when the probe lengths vary (the ends are then not sorted and
`#(start <= qe)` takes the start-keyed search + forward walk instead of the interleaved dual search)"""
import sys
import torch
sys.path.insert(0, ".")
import bench as B
import bio_ranges_b200 as brg

dev = torch.device("cuda:0")
ctx = brg.default_context(0)
b = B.gen_intervals(torch, 10_000_000, 42, 100, 10_000, dev)
ix = ctx.build_dev(*b, len(B.HG38), brg.WANT_COUNT)
for name, lo, hi in [("fixed 150 bp", 150, 150), ("U[100, 200]", 100, 200), ("U[1, 300]", 1, 300), ("U[1, 3000]", 1, 3000)]:
    p = B.order_rows(torch, *B.gen_intervals(torch, 125_000_000, 43, lo, hi, dev), "sorted", 43)
    out = torch.empty(p[1].numel(), dtype=torch.int64, device=dev)
    for _ in range(3):
        ix.count_overlaps_dev(*p, out)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        ix.count_overlaps_dev(*p, out)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    print("probe length %-13s %.4f ms  %.1f %% of the HBM roofline" % (name, ms, 100 * 2.58e9 / (ms * 1e-3) / 6554.9e9))
    del p, out
