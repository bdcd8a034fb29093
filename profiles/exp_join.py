"""pair join: device-resident time of the one-launch join (brg_pairs_join_dev) next to the three-launch chunked form on the dense
(10M x 10M, configs[1]) and the sparse (10M x 1M) workload, sorted and shuffled.  --check: parity of both against the oracle's
order-independent checksum.  BRG_LIB selects a variant library (profiles/build_variant.sh)."""
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
cases = (("dense 10Mx10M", 1.0, "sorted"), ("sparse 10Mx1M", 0.1, "sorted"), ("dense 10Mx10M", 1.0, "shuffled"))
if "--sorted-only" in sys.argv:
    cases = cases[:2]
def t(fn, reps=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
for label, nb_scale, order in cases:
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
    tot = torch.zeros(1, dtype=torch.int64, device=dev)
    alg = n * 12 + total * 8 + b[1].numel() * 12
    rec = {"case": label, "order": order, "lib": os.path.basename(os.environ.get("BRG_LIB", "libbioranges.so")), "pairs": total}
    want = None
    if check:
        oi = O.OracleIndex(b[0].cpu().numpy().astype(np.uint32), b[1].cpu().numpy(), b[2].cpu().numpy(), len(bench.HG38))
        want = oi.overlap_checksum(pk.cpu().numpy().astype(np.uint32), ps.cpu().numpy(), pe.cpu().numpy(), threads=O.max_threads())
    if "--no-chunked" not in sys.argv:
        t1 = t(lambda: ix.pairs_chunk_offsets_dev(pk, ps, pe, offsets, ranges))
        t2 = t(lambda: ix.pairs_fill_dev(ps, offsets, ranges, bi, pi))
        rec.update(chunked_ms=round(t1 + t2, 4), chunked_frac=round(alg / ((t1 + t2) * 1e-3) / 1e9 / 6554.9, 3))
        if check:
            rec["chunked_parity"] = O.pair_checksum(bi.cpu().numpy().view(np.uint32), pi.cpu().numpy().view(np.uint32)) == want
    bi.fill_(-1); pi.fill_(-1)
    tj = t(lambda: ix.pairs_join_dev(pk, ps, pe, bi, pi, tot))
    rec.update(join_ms=round(tj, 4), join_frac=round(alg / (tj * 1e-3) / 1e9 / 6554.9, 3), join_total_ok=int(tot.item()) == total)
    if check:
        rec["join_parity"] = O.pair_checksum(bi.cpu().numpy().view(np.uint32), pi.cpu().numpy().view(np.uint32)) == want
    if "--no-ordered" not in sys.argv:
        bi.fill_(-1); pi.fill_(-1)
        to = t(lambda: ix.pairs_join_dev(pk, ps, pe, bi, pi, tot, ordered=True))
        rec.update(join_ordered_ms=round(to, 4), join_ordered_total_ok=int(tot.item()) == total)
        if check:
            rec["join_ordered_parity"] = O.pair_checksum(bi.cpu().numpy().view(np.uint32), pi.cpu().numpy().view(np.uint32)) == want
    if "--prof" in sys.argv:  # a library built with -DBRG_FJ_PROF: per-phase clock sums of one launch, as shares
        import ctypes as C
        L = brg._cabi.lib()
        if hasattr(L, "brg_debug_join_prof"):
            buf = (C.c_ulonglong * 16)()
            L.brg_debug_join_prof(None, 1)
            ix.pairs_join_dev(pk, ps, pe, bi, pi, tot); torch.cuda.synchronize()
            L.brg_debug_join_prof(buf, 0)
            names = ["load", "bounds", "stage_window", "ranks", "map", "test", "lookback", "copy", "slow"]
            tot_c = sum(buf[i] for i in range(9)) or 1
            rec["prof_share"] = {names[i]: round(buf[i] / tot_c, 3) for i in range(9)}
            rec["prof_cycles_per_chunk"] = round(tot_c / max(1, buf[9] + buf[10]))
            rec["chunks_fast_slow"] = [int(buf[9]), int(buf[10])]
            nch = max(1, buf[9] + buf[10])
            rec["lookback_detail"] = {"group_level_cycles_per_chunk": round(buf[11] / nch), "in_group_spins_per_chunk": round(buf[12] / nch, 2),
                                      "group_spins_per_chunk": round(buf[13] / nch, 2), "group_walk_steps_per_chunk": round(buf[14] / nch, 2),
                                      "publish_cycles_per_chunk": round(buf[15] / nch)}
    print(json.dumps(rec), flush=True)
    ix.close()
    del bi, pi, offsets, ranges
    torch.cuda.empty_cache()
