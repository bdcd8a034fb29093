"""count_overlaps on a SHUFFLED probe table (125M x 10M, 24 contigs): device-resident time of the per-row path.
BRG_CB_TARGET = elements of each array per count bin (0 = no bins: the compact rank index of round 1)."""
import os, sys, json
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench
import bio_ranges_b200 as brg
from oracle import oracle as O

ctx = brg.Context(0)
dev = "cuda:0"
b, p = bench.make_workload(torch, "count", dev, 0, 1, "shuffled", 1.0)
nb = int(os.environ.get("BUILD_ROWS", "0"))
if nb:
    b = tuple(x[:nb].contiguous() for x in b)
ix = ctx.build_dev(*b, len(bench.HG38), brg.WANT_COUNT)
pk, ps, pe = p
n = ps.numel()
out = torch.empty(n, dtype=torch.int64, device=dev)
for _ in range(3):
    ix.count_overlaps_dev(pk, ps, pe, out)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5):
    ix.count_overlaps_dev(pk, ps, pe, out)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 5
m = 4_000_000
oi = O.OracleIndex(b[0].cpu().numpy().astype(np.uint32), b[1].cpu().numpy(), b[2].cpu().numpy(), len(bench.HG38))
want = oi.count_overlaps(pk[:m].cpu().numpy().astype(np.uint32), ps[:m].cpu().numpy(), pe[:m].cpu().numpy(), threads=O.max_threads())
ok = bool((out[:m].cpu().numpy() == want).all())
print(json.dumps({"build_rows": b[1].numel(), "cb_target": os.environ.get("BRG_CB_TARGET", "3"), "ms": round(ms, 4), "roofline_frac": round((n * 20 + 8 * b[1].numel()) / (ms * 1e-3) / 1e9 / 6554.9, 3),
                  "index_mb": round(ix.device_bytes / 1e6, 1), "parity_4m": ok}))
