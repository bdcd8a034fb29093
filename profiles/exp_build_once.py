"""one warm index build (for a launch list under ncu): python profiles/exp_build_once.py [count|coverage|nearest|pairs]"""
import sys
import torch
sys.path.insert(0, ".")
import bench as B
import bio_ranges_b200 as brg
dev = torch.device("cuda:0")
ctx = brg.default_context(0)
b, p = B.make_workload(torch, "count", dev, 0, 1, "sorted", 1.0)
want = {"count": brg.WANT_COUNT, "coverage": brg.WANT_COVERAGE, "nearest": brg.WANT_NEAREST, "pairs": brg.WANT_PAIRS}[sys.argv[1] if len(sys.argv) > 1 else "count"]
for i in range(2):
    ix = ctx.build_dev(*b, len(B.HG38), want)
    torch.cuda.synchronize()
    ix.close()
