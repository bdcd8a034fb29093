"""pair join: device-resident time of the two passes on the dense (10M x 10M) and the sparse (10M x 1M) workload,
sorted and shuffled; BRG_PAIRS_V2=0 selects the round-1 kernels.  Parity by the order-independent checksum."""
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
check = "--check" in sys.argv
for label, nb_scale, order in (("dense 10Mx10M", 1.0, "sorted"), ("sparse 10Mx1M", 0.1, "sorted"), ("dense 10Mx10M", 1.0, "shuffled")):
    b, p = bench.make_workload(torch, "pairs", dev, 0, 1, order, 1.0)
    if nb_scale != 1.0:
        nb = int(b[1].numel() * nb_scale)
        b = tuple(x[:nb].contiguous() for x in b)
    ix = ctx.build_dev(*b, len(bench.HG38), brg.WANT_PAIRS)
    pk, ps, pe = p
    n = ps.numel()
    offsets = torch.empty((n + 255) // 256 + 1, dtype=torch.int64, device=dev)
    ranges = torch.empty(n, dtype=torch.int64, device=dev)
    ix.pairs_chunk_offsets_dev(pk, ps, pe, offsets, ranges)
    torch.cuda.synchronize()
    total = int(offsets[-1].item())
    bi = torch.empty(max(total, 1), dtype=torch.int32, device=dev)
    pi = torch.empty(max(total, 1), dtype=torch.int32, device=dev)
    def t(fn, reps=10):
        for _ in range(3): fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps): fn()
        e1.record(); torch.cuda.synchronize()
        return e0.elapsed_time(e1) / reps
    t1 = t(lambda: ix.pairs_chunk_offsets_dev(pk, ps, pe, offsets, ranges))
    t2 = t(lambda: ix.pairs_fill_dev(ps, offsets, ranges, bi, pi))
    alg = n * 12 + total * 8 + b[1].numel() * 12
    ok = None
    if check:
        oi = O.OracleIndex(b[0].cpu().numpy().astype(np.uint32), b[1].cpu().numpy(), b[2].cpu().numpy(), len(bench.HG38))
        want = oi.overlap_checksum(pk.cpu().numpy().astype(np.uint32), ps.cpu().numpy(), pe.cpu().numpy(), threads=O.max_threads())
        ok = O.pair_checksum(bi.cpu().numpy().view(np.uint32), pi.cpu().numpy().view(np.uint32)) == want
    print(json.dumps({"case": label, "order": order, "v2": os.environ.get("BRG_PAIRS_V2", "1"), "pairs": total, "pass1_ms": round(t1, 4), "pass2_ms": round(t2, 4),
                      "total_ms": round(t1 + t2, 4), "roofline_frac": round(alg / ((t1 + t2) * 1e-3) / 1e9 / 6554.9, 3), "parity": ok}))
    ix.close()
