"""configs[0] (count_overlaps 1M reads x 100k targets, 1 contig; sorted and shuffled) and a sparse pair join (10M x 1M),
device-resident, CUDA events"""
import sys
import torch
sys.path.insert(0, ".")
import bench as B
import bio_ranges_b200 as brg

dev = torch.device("cuda:0")
ctx = brg.default_context(0)


def timed(fn, reps=20):
    for _ in range(3):
        fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


g = torch.Generator(device=dev).manual_seed(42)
nb, npb, L = 100_000, 1_000_000, 2**28
bs = torch.randint(0, L, (nb,), generator=g, device=dev, dtype=torch.int32)
be = bs + torch.randint(100, 10_001, (nb,), generator=g, device=dev, dtype=torch.int32)
bk = torch.zeros(nb, dtype=torch.int32, device=dev)
ps = torch.randint(0, L, (npb,), generator=g, device=dev, dtype=torch.int32)
ix = ctx.build_dev(bk, bs, be, 1, brg.WANT_COUNT)
out = torch.empty(npb, dtype=torch.int64, device=dev)
for order in ("sorted", "shuffled"):
    s = torch.sort(ps).values if order == "sorted" else ps
    e = s + 150
    k = torch.zeros(npb, dtype=torch.int32, device=dev)
    ms = timed(lambda: ix.count_overlaps_dev(k, s, e, out))
    print("configs[0] count_overlaps 1M x 100k, 1 contig, %-8s %.1f us per call, %.2e probes/s" % (order, ms * 1e3, npb / (ms * 1e-3)))

b = B.gen_intervals(torch, 1_000_000, 42, 1, 1550, dev)
p = B.order_rows(torch, *B.gen_intervals(torch, 10_000_000, 43, 1, 1550, dev), "sorted", 43)
ixp = ctx.build_dev(*b, len(B.HG38), brg.WANT_PAIRS)
n = p[1].numel()
n_chunks = (n + 255) // 256
cb = torch.empty(n_chunks + 1, dtype=torch.int64, device=dev)
rr = torch.empty(n, dtype=torch.int64, device=dev)
ixp.pairs_chunk_offsets_dev(*p, cb, rr)
total = int(cb[-1].item())
bi = torch.empty(total, dtype=torch.int32, device=dev)
pi = torch.empty(total, dtype=torch.int32, device=dev)


def step():
    ixp.pairs_chunk_offsets_dev(*p, cb, rr)
    ixp.pairs_fill_dev(p[1], cb, rr, bi, pi)


ms = timed(step, 10)
print("sparse pair join 10M x 1M: %d pairs, %.3f ms, %.2e pairs/s, %.1f %% of the HBM roofline" % (
    total, ms, total / (ms * 1e-3), 100 * (n * 12 + total * 8 + 1_000_000 * 12) / (ms * 1e-3) / 6554.9e9))
