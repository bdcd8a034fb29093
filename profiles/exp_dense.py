"""ad-hoc experiment: dense-window regime (10M probes x 10M build) -- tiled kernels (dense_ranks) vs one-row-per-thread
kernels, for count / coverage / nearest.  usage: [BRG_TILED=...] python profiles/exp_dense.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench as B
import bio_ranges_b200 as brg

dev = "cuda:0"
ctx = brg.Context(0)
b, p = B.make_workload(torch, "pairs", dev, 0, 1, "sorted")
pk, ps, pe = p
n = ps.numel()
ix = ctx.build_dev(*b, len(B.HG38), brg.WANT_ALL)
out = torch.empty(n, dtype=torch.int64, device=dev)
left = torch.empty(n, dtype=torch.int32, device=dev)
bits = torch.empty((n + 31) // 32, dtype=torch.int32, device=dev)
dd = torch.empty(n, dtype=torch.int64, device=dev)
ops = {"count": lambda: ix.count_overlaps_dev(pk, ps, pe, out), "coverage": lambda: ix.coverage_dev(pk, ps, pe, out),
       "nearest": lambda: ix.nearest_k1_dev(pk, ps, pe, left, bits, dd)}
for name, f in ops.items():
    for _ in range(3):
        f()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        f()
    e1.record()
    torch.cuda.synchronize()
    print(os.environ.get("BRG_TILED", "default"), name, "ms/step", e0.elapsed_time(e1) / 10)
