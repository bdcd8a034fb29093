"""index build time (device-resident input), first call vs warm calls, per want mask"""
import sys, time
import torch
sys.path.insert(0, ".")
import bench as B
import bio_ranges_b200 as brg

dev = torch.device("cuda:0")
ctx = brg.default_context(0)
b, p = B.make_workload(torch, "count", dev, 0, 1, "sorted", 1.0)
for name, want in [("count", brg.WANT_COUNT), ("coverage", brg.WANT_COVERAGE), ("nearest", brg.WANT_NEAREST), ("pairs", brg.WANT_PAIRS)]:
    ts = []
    for i in range(4):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        ix = ctx.build_dev(*b, len(B.HG38), want)
        torch.cuda.synchronize(); ts.append(time.perf_counter() - t0)
        ix.close()
    print(name, "build of", b[1].numel(), "rows: first %.1f ms, warm %s ms" % (ts[0] * 1e3, [round(t * 1e3, 1) for t in ts[1:]]))
