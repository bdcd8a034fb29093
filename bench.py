#!/usr/bin/env python
"""bench.py -- throughput of the interval-overlap hot path on N B200s (one process per GPU).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--workload count|pairs|coverage|nearest] [--order sorted|shuffled]
  python bench.py --impl reference ...      # the reference's CPU algorithm (oracle port) on the host cores

Headline workload (BASELINE.json north_star / configs[4]): count_overlaps, 125M probe intervals per GPU x 10M build
intervals over 24 hg38-like contigs (N=8 => 1B x 10M), weak scaling: the build index is built on rank 0 and replicated
with an NCCL broadcast, probe rows are sharded by contiguous row range, no collective on the data path.

One JSON line on rank 0:
  value     probes/s, whole job, inputs resident in HBM, K timed steps bracketed by barrier + synchronize, max over ranks
  e2e       same metric through the host-buffer C-ABI call (brg_count_overlaps: per-row key id column, start, end in PINNED
            host buffers -> int64 counts in a pinned host buffer); H2D + the library's per-chunk key-run scan + kernels + D2H
            inside the timed region.  e2e_pageable: the same call on ordinary heap buffers (what arrow-rs allocates)
  roofline  dominant kernel vs the measured HBM copy bandwidth (MEASURED_PEAKS.json); roofline_with_build = index build + one
            probe step (SURVEY 8d's build-amortised figure)
  parity    EVERY rank compares ALL rows of its device-resident output and of its end-to-end output with the oracle
            (pairs: per-probe counts + an order-independent checksum of the pair set); min over ranks
  extras    the other section-8 workloads (coverage, nearest, pairs; sorted and shuffled), each with its own parity, at every N
  cpu_baseline  the oracle port (reference algorithm restated in C + OpenMP) on a bounded sample, rank 0, N=1
"""
import argparse
import json
import os
import subprocess
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

# hg38 primary contig lengths (chr1..22, X, Y), used only as relative sizes of the synthetic contigs
HG38 = [248956422, 242193529, 198295559, 190214555, 181538259, 170805979, 159345973, 145138636, 138394717, 133797422,
        135086622, 133275309, 114364328, 107043718, 101991189, 90338345, 83257441, 80373285, 58617616, 64444167,
        46709983, 50818468, 156040895, 57227415]

WORKLOADS = {
    # name: (probes per GPU, build rows, description)
    "count": (125_000_000, 10_000_000, "count_overlaps 125M probe/GPU x 10M build, 24 contigs (configs[4] shard; N=8 => 1B x 10M)"),
    "pairs": (10_000_000, 10_000_000, "overlap pairs (IntervalJoinExec) 10M x 10M, 24 contigs, ~5 pairs/probe (configs[1])"),
    "coverage": (100_000_000, 1_000_000, "coverage 100M reads (150 bp) x 1M exon-like targets, 24 contigs (configs[2])"),
    "nearest": (50_000_000, 5_000_000, "nearest k=1 50M queries x 5M sparse targets, 24 contigs (configs[3])"),
}
KERNELS = {"count": "count_kernel_tiled", "coverage": "coverage_kernel_tiled", "nearest": "nearest_k1_kernel_win",
           "pairs": "pairs_join_kernel"}
API = {"count": "brg_count_overlaps", "coverage": "brg_coverage", "nearest": "brg_nearest_k1", "pairs": "brg_overlap_pairs"}


def hbm_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def ncu_traffic(kernel, workload, order, scale):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of the dominant kernel, from the committed
    `ncu --set full` capture of this same workload (profiles/traffic.json); None when the run differs from it."""
    p = os.path.join(ROOT, "profiles", "traffic.json")
    if not os.path.exists(p) or scale != 1.0:
        return None
    rec = json.load(open(p)).get(f"{workload}:{order}")
    return rec["dram_bytes_per_launch"] if rec and rec.get("kernel") == kernel else None


# ----------------------------------------------------------------------------------------------------------
# synthetic data (torch Philox on the device; seeds 42 build / 43 probe as in SURVEY 8d)
# ----------------------------------------------------------------------------------------------------------
def gen_intervals(torch, n, seed, len_lo, len_hi, device, quantile=(0.0, 1.0), lognormal=None, clustered=False):
    """contig ~ proportional to length, start ~ U over the contig, length ~ U[len_lo, len_hi] (or lognormal).
    `quantile` restricts positions to a slice of the concatenated genome (contiguous shard of a sorted table)."""
    g = torch.Generator(device=device).manual_seed(seed)
    lens = torch.tensor(HG38, dtype=torch.float64, device=device)
    cum = torch.cumsum(lens, 0)
    total = float(cum[-1])
    u = torch.rand(n, generator=g, device=device, dtype=torch.float64)
    if clustered:  # exon-like: positions drawn around 200k cluster centres
        centres = torch.rand(200_000, generator=g, device=device, dtype=torch.float64)
        pick = torch.randint(0, 200_000, (n,), generator=g, device=device)
        u = (centres[pick] + (u - 0.5) * 2e-5).clamp_(0.0, 1.0 - 1e-12)
    pos = (quantile[0] + u * (quantile[1] - quantile[0])) * total
    key = torch.searchsorted(cum, pos, right=True).clamp_(max=len(HG38) - 1)
    base = cum[key] - lens[key]
    start = (pos - base).to(torch.int32)
    if lognormal is not None:
        med, sigma, lo, hi = lognormal
        z = torch.randn(n, generator=g, device=device, dtype=torch.float32)
        length = (med * torch.exp(sigma * z)).clamp_(lo, hi).to(torch.int32)
    else:
        length = torch.randint(len_lo, len_hi + 1, (n,), generator=g, device=device, dtype=torch.int32)
    end = start + length
    return key.to(torch.int32), start, end


def order_rows(torch, key, start, end, order, seed):
    if order == "sorted":
        comp = key.to(torch.int64) * (1 << 32) + start.to(torch.int64)
        idx = torch.argsort(comp)
    else:
        g = torch.Generator(device=key.device).manual_seed(seed + 1000)
        idx = torch.randperm(key.numel(), generator=g, device=key.device)
    return key[idx].contiguous(), start[idx].contiguous(), end[idx].contiguous()


def make_workload(torch, name, device, rank, world, order, scale=1.0):
    n_probe, n_build, _ = WORKLOADS[name]
    n_probe = max(1024, int(n_probe * scale))
    n_build = max(1024, int(n_build * scale))
    q = (rank / world, (rank + 1) / world) if order == "sorted" else (0.0, 1.0)
    if name == "count":
        b = gen_intervals(torch, n_build, 42, 100, 10_000, device)
        p = gen_intervals(torch, n_probe, 43 + rank, 150, 150, device, quantile=q)
    elif name == "pairs":
        b = gen_intervals(torch, n_build, 42, 1, 1550, device)
        p = gen_intervals(torch, n_probe, 43 + rank, 1, 1550, device, quantile=q)
    elif name == "coverage":
        b = gen_intervals(torch, n_build, 42, 0, 0, device, lognormal=(150.0, 1.0, 10, 10_000), clustered=True)
        p = gen_intervals(torch, n_probe, 43 + rank, 150, 150, device, quantile=q)
    else:
        b = gen_intervals(torch, n_build, 42, 50, 500, device)
        p = gen_intervals(torch, n_probe, 43 + rank, 1, 300, device, quantile=q)
    p = order_rows(torch, *p, order, 43 + rank)
    return b, p


# ----------------------------------------------------------------------------------------------------------
# clocks
# ----------------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, device_index):
        self.dev = device_index
        self.path = f"/tmp/brg_clocks_{os.getpid()}.csv"
        self.p = None

    def start(self):
        try:
            self.p = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "20",
                                       "-i", str(self.dev)], stdout=open(self.path, "w"), stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        if self.p is None:
            return out
        time.sleep(0.15)
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        sm, mx, reasons = [], [], set()
        for line in open(self.path):
            f = [x.strip() for x in line.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        if sm:
            out = {"sm_mhz": float(np.median(sm)), "sm_max_mhz": max(mx), "reasons": sorted(reasons), "samples": len(sm)}
        try:
            os.remove(self.path)
        except OSError:
            pass
        return out


# ----------------------------------------------------------------------------------------------------------
# CPU arm: the reference's algorithm (oracle port) on the host cores
# ----------------------------------------------------------------------------------------------------------
def cpu_run(name, b, p, sample_rows, steps, warmup):
    """returns (rows/s, cores, sample description, build seconds).  b/p are numpy column tuples."""
    from oracle import oracle as O

    cores = O.max_threads()
    bk, bs, be = b
    pk, ps, pe = (x[:sample_rows] for x in p)
    t0 = time.perf_counter()
    oi = O.OracleIndex(bk, bs, be, len(HG38))
    t_build = time.perf_counter() - t0
    if name == "count":
        fn = lambda: oi.count_overlaps(pk, ps, pe, threads=cores)
    elif name == "coverage":
        fn = lambda: oi.coverage(pk, ps, pe, threads=cores)
    elif name == "nearest":
        fn = lambda: oi.nearest_k1(pk, ps, pe, threads=cores)
    else:
        # the join as the reference runs it on `cores` partitions: every thread materialises the (build, probe) index vectors of
        # its probe-row range (oracle: brgo_overlap_pairs_partitions, IJ:1616-1653)
        fn = lambda: oi.overlap_pairs_partitions(pk, ps, pe, threads=cores)
    for _ in range(warmup):
        fn()
    t0 = time.perf_counter()
    for _ in range(steps):
        res = fn()
    dt = (time.perf_counter() - t0) / steps
    units = len(ps)
    if name == "pairs":
        units = int(res)
    return units / dt, cores, (f"first {len(ps)} probe rows of the workload (coordinate-sorted: mostly the first contigs), {steps} timed "
                               f"passes, index build {t_build:.2f}s excluded; C port of the reference algorithm, not the Rust crate"), dt


def to_np(t):
    return t.cpu().numpy()


def split_affinity(local_rank, world):
    """N processes on one box: give each rank its own slice of the cores this job may use (the NVML "ideal" affinity of every
    GPU of a single-NUMA-node box is ALL cores, which leaves every rank's host threads on the same cores)."""
    try:
        cores = sorted(os.sched_getaffinity(0))
        per = len(cores) // world
        if per < 1:
            return len(cores)
        mine = cores[local_rank * per:(local_rank + 1) * per]
        os.sched_setaffinity(0, mine)
        return len(mine)
    except Exception:
        return 0


class Harness:
    """process-wide state shared by the headline and the extras"""

    def __init__(self, torch, dist, brg, rank, local_rank, world, device, args):
        self.torch, self.dist, self.brg = torch, dist, brg
        self.rank, self.local_rank, self.world, self.device, self.args = rank, local_rank, world, device, args
        self.ctx = brg.Context(local_rank)
        self.lib = brg._cabi.lib()
        self.peak, self.peak_src = hbm_peak()
        self._oracle_cache = {}

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def reduce(self, v, op):
        if self.world == 1:
            return v
        t = self.torch.tensor([v], dtype=self.torch.float64, device=self.device)
        self.dist.all_reduce(t, op={"max": self.dist.ReduceOp.MAX, "sum": self.dist.ReduceOp.SUM, "min": self.dist.ReduceOp.MIN}[op])
        return float(t.item())

    def oracle_threads(self):
        from oracle import oracle as O

        return max(1, O.max_threads())

    def oracle_index(self, name, b):
        """the checker's index over the build side of workload `name` (one per workload and process)"""
        from oracle import oracle as O

        if name not in self._oracle_cache:
            self._oracle_cache.clear()  # one at a time: 10M-row indexes are not small
            self._oracle_cache[name] = O.OracleIndex(to_np(b[0]).astype(np.uint32), to_np(b[1]), to_np(b[2]), len(HG38))
        return self._oracle_cache[name]


def check_fixed(h, name, oi, pnp, got):
    """ALL rows of a fixed-size result against the oracle; got = numpy arrays as the C ABI returns them"""
    thr = h.oracle_threads()
    npk, nps, npe = pnp
    if name == "count":
        return bool((got[0] == oi.count_overlaps(npk, nps, npe, threads=thr)).all())
    if name == "coverage":
        return bool((got[0] == oi.coverage(npk, nps, npe, threads=thr)).all())
    left, bits, dist_ = got
    ol, ov, od = oi.nearest_k1(npk, nps, npe, threads=thr)
    valid = np.unpackbits(bits.view(np.uint8), bitorder="little")[: len(nps)].astype(bool)
    return bool((valid == ov).all() and (left.view(np.uint32)[ov] == ol[ov]).all() and (dist_[ov] == od[ov]).all())


def check_pairs(h, oi, pnp, bi, pi):
    """per-probe counts + the order-independent checksum (sum h, sum h^2, xor h) of the whole pair set"""
    from oracle import oracle as O

    npk, nps, npe = pnp
    want, cnt = oi.overlap_checksum(npk, nps, npe, row_base=0, threads=h.oracle_threads(), counts=True)
    bi, pi = np.asarray(bi).view(np.uint32), np.asarray(pi).view(np.uint32)
    if len(bi) != want[3]:
        return False
    got_cnt = np.bincount(pi, minlength=len(nps)).astype(np.uint64)
    return bool((got_cnt == cnt).all() and O.pair_checksum(bi, pi) == want)


def run_workload(h, name, order, steps, warmup, want_e2e, want_pageable=False, want_build_fig=False, sampler=None):
    """one workload on this rank's shard: build + replicate, device-resident timed steps, end-to-end steps, parity"""
    torch, dist, brg = h.torch, h.dist, h.brg
    rank, world, device, args = h.rank, h.world, h.device, h.args
    b, p = make_workload(torch, name, device, rank, world, order, args.scale)
    want = {"count": brg.WANT_COUNT, "pairs": brg.WANT_PAIRS, "coverage": brg.WANT_COVERAGE, "nearest": brg.WANT_NEAREST}[name]
    nb = b[1].numel()
    res = {"workload": WORKLOADS[name][2], "probe_order": order}

    # ---- build on rank 0, replicate the index slab (it carries its own header) with ONE NCCL broadcast ----------------------
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    ix = None
    if rank == 0:
        ix = h.ctx.build_dev(*b, len(HG38), want)
        torch.cuda.synchronize()
    build_first_s = time.perf_counter() - t0
    build = {"seconds_rank0_first_call": build_first_s}
    if world > 1:
        from bio_ranges_b200.sharding import replicate_index

        times = []
        for rep in range(2):  # first call: NCCL channel set-up + cudaMalloc of the replica slab; second call: the transfer itself
            h.barrier()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            t0 = time.perf_counter()
            e0.record()
            rep_ix = replicate_index(h.ctx, ix if rank == 0 else None, dist, rank, world, device)
            e1.record()
            torch.cuda.synchronize()
            times.append((h.reduce(e0.elapsed_time(e1), "max"), h.reduce((time.perf_counter() - t0) * 1e3, "max")))
            if rep == 0 and rank != 0:
                rep_ix.close()
        ix = rep_ix
        build["broadcast_ms_first_call"] = {"device": times[0][0], "wall": times[0][1]}
        build["broadcast_ms"] = {"device": times[1][0], "wall": times[1][1]}
        build["broadcast_gbs"] = ix.device_bytes / (times[1][0] * 1e-3) / 1e9
    build["index_bytes"] = ix.device_bytes
    pk, ps, pe = p
    n = ps.numel()

    # ---- device-resident step ------------------------------------------------------------------------------
    total = None
    if name in ("count", "coverage"):
        out = torch.empty(n, dtype=torch.int64, device=device)
        fn = (ix.count_overlaps_dev if name == "count" else ix.coverage_dev)
        step = lambda: fn(pk, ps, pe, out)
        units_per_step = n
        alg_bytes = n * 20 + nb * 8
        outs = (out,)
    elif name == "nearest":
        left = torch.empty(n, dtype=torch.int32, device=device)
        bits = torch.zeros((n + 31) // 32, dtype=torch.int32, device=device)
        dd = torch.empty(n, dtype=torch.int64, device=device)
        step = lambda: ix.nearest_k1_dev(pk, ps, pe, left, bits, dd)
        units_per_step = n
        alg_bytes = n * 28 + n // 8 + nb * 12
        outs = (left, bits, dd)
    else:
        # sorted tables: the join in ONE launch, any-order output slots (brg_pairs_join_dev); shuffled tables: the three-launch
        # chunked form, whose one-row-per-thread kernels are the faster ones there (profiles/r2_join_experiments.txt)
        fused = os.environ.get("BRG_PAIRS_FUSED", "1") != "0" and order == "sorted"
        offsets = torch.empty((n + 255) // 256 + 1, dtype=torch.int64, device=device)  # per 256-row chunk
        ranges = torch.empty(n, dtype=torch.int64, device=device)  # per-row candidate ranges (pass 1 -> pass 2 of the chunked form)
        ix.pairs_chunk_offsets_dev(pk, ps, pe, offsets, ranges)
        torch.cuda.synchronize()
        total = int(offsets[-1].item())
        bi = torch.empty(max(total, 1), dtype=torch.int32, device=device)
        pi = torch.empty(max(total, 1), dtype=torch.int32, device=device)
        tot_dev = torch.zeros(1, dtype=torch.int64, device=device)
        if fused:
            del ranges, offsets
            ranges = offsets = None

            def step():  # ONE launch: ranges -> count -> decoupled look-back -> emit (brg_pairs_join_dev)
                ix.pairs_join_dev(pk, ps, pe, bi, pi, tot_dev)
        else:

            def step():  # count -> scan -> emit (three-launch chunked form)
                ix.pairs_chunk_offsets_dev(pk, ps, pe, offsets, ranges)
                ix.pairs_fill_dev(ps, offsets, ranges, bi, pi)

        res["pairs_path"] = "brg_pairs_join_dev (one launch)" if fused else "brg_pairs_chunk_offsets_dev + brg_pairs_fill_dev"
        units_per_step = total
        alg_bytes = n * 12 + total * 8 + nb * 12
        outs = (bi, pi)
    for _ in range(warmup):
        step()
    h.barrier()
    if sampler is not None and rank == 0:
        sampler.start()
    l0 = h.lib.brg_kernel_launches()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(steps + 1)]
    h.barrier()
    ev[0].record()
    for i in range(steps):
        step()
        ev[i + 1].record()
    torch.cuda.synchronize()
    h.barrier()
    total_ms = h.reduce(ev[0].elapsed_time(ev[-1]), "max")
    launches = h.lib.brg_kernel_launches() - l0
    per_step_ms = [ev[i].elapsed_time(ev[i + 1]) for i in range(steps)]
    ms_per_step = total_ms / steps
    res["value"] = h.reduce(units_per_step, "sum") / (ms_per_step * 1e-3)
    res["unit"] = "pairs/s" if name == "pairs" else "probes/s"
    res["ms_per_step"] = ms_per_step
    res["gpu_launches"] = int(launches)
    kern_ms = float(np.mean(per_step_ms))
    achieved = alg_bytes / (kern_ms * 1e-3) / 1e9
    kernel_name = KERNELS[name] if (name != "pairs" or res.get("pairs_path", "").startswith("brg_pairs_join_dev")) else \
        "pairs_ranges_todo_kernel+scan+pairs_emit_kernel_flat"
    if order == "shuffled" and name != "pairs":  # unsorted rows: the tiled kernel hands every chunk to the one-row-per-thread kernel
        kernel_name = {"count": "count_kernel_todo", "coverage": "coverage_kernel_todo", "nearest": "nearest_k1_kernel_todo"}[name]
    res["roofline"] = {"bound": "hbm", "achieved": achieved, "peak": h.peak, "unit": "GB/s", "frac": achieved / h.peak,
                       "traffic": ncu_traffic(kernel_name, name, order, args.scale), "peak_source": h.peak_src, "kernel": kernel_name,
                       "algorithmic_bytes_per_launch": alg_bytes, "launch_ms": kern_ms}
    if name == "pairs":
        res["pairs_per_probe"] = total / n
        if res["pairs_path"].startswith("brg_pairs_join_dev") and rank == 0:
            # the same join with probe-major output throughout, for the record: the one-launch kernel with its look-back
            # (BRG_JOIN_ORDERED) and the three-launch chunked form that brg_overlap_pairs(BRG_PAIRS_INNER) takes
            def timed(fn, reps=max(3, steps)):
                for _ in range(3):
                    fn()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for _ in range(reps):
                    fn()
                e1.record()
                torch.cuda.synchronize()
                return e0.elapsed_time(e1) / reps

            off2 = torch.empty((n + 255) // 256 + 1, dtype=torch.int64, device=device)
            rng2 = torch.empty(n, dtype=torch.int64, device=device)
            t_ord = timed(lambda: ix.pairs_join_dev(pk, ps, pe, bi, pi, tot_dev, ordered=True))
            t_chk = timed(lambda: (ix.pairs_chunk_offsets_dev(pk, ps, pe, off2, rng2), ix.pairs_fill_dev(ps, off2, rng2, bi, pi)))
            res["pairs_ordered_output"] = {"one_launch_look_back_ms": t_ord, "three_launch_chunked_ms": t_chk,
                                           "what": "same join, output probe-major throughout (the default lets the 128-row chunks land in the order they finish)"}
            del off2, rng2
            ix.pairs_join_dev(pk, ps, pe, bi, pi, tot_dev)  # (the parity check below reads the any-order result that was timed)
            torch.cuda.synchronize()

    # ---- parity of the DEVICE-RESIDENT result that produced `value`: all rows, this rank ----------------------------------
    parity = {}
    oi = None
    pnp = None
    if not args.no_parity:
        oi = h.oracle_index(name, b)
        pnp = (to_np(pk).astype(np.uint32), to_np(ps), to_np(pe))
        if name == "pairs":
            ok = check_pairs(h, oi, pnp, to_np(bi[:total]), to_np(pi[:total]))
            if res.get("pairs_path", "").startswith("brg_pairs_join_dev"):
                ok = ok and int(tot_dev.item()) == total  # the one-launch join's own count
        else:
            ok = check_fixed(h, name, oi, pnp, tuple(to_np(o) for o in outs))
        parity["device_resident"] = bool(h.reduce(1.0 if ok else 0.0, "min") == 1.0)

    # ---- index build amortised into one probe step: a second build with the scratch pool warm ------------------
    if want_build_fig and rank == 0:
        h.ctx.build_dev(*b, len(HG38), want).close()  # (leaves a slab of this size in the scratch pool, as a previous query would)
        torch.cuda.synchronize()
        e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
        e0.record()
        ix2 = h.ctx.build_dev(*b, len(HG38), want)
        e1.record()
        step()
        e2.record()
        torch.cuda.synchronize()
        ix2.close()
        bms, sms = e0.elapsed_time(e1), e1.elapsed_time(e2)
        build["ms_warm"] = bms
        res["roofline_with_build"] = {"build_ms": bms, "step_ms": sms, "frac": alg_bytes / ((bms + sms) * 1e-3) / 1e9 / h.peak,
                                      "value": units_per_step / ((bms + sms) * 1e-3), "unit": res["unit"],
                                      "what": "one index build (device-resident build columns, scratch pool warm) + one probe step"}
    res["build"] = build
    del outs
    if name == "pairs":
        del bi, pi, offsets, ranges
    torch.cuda.empty_cache()

    # ---- end to end through the host-buffer C-ABI call: host inputs -> H2D -> kernels -> D2H ---------------------------------
    if want_e2e:
        def pinned(t):
            x = torch.empty(t.numel(), dtype=t.dtype).pin_memory()
            x.copy_(t)
            return x

        h_pk, h_ps, h_pe = pinned(pk), pinned(ps), pinned(pe)
        npk, nps, npe = h_pk.numpy().view(np.uint32), h_ps.numpy(), h_pe.numpy()
        e2e_steps = max(3, min(steps, 5))

        def make_step(k_, s_, e_, pin):
            alloc = (lambda m, dt: torch.empty(m, dtype=dt).pin_memory().numpy()) if pin else (lambda m, dt: torch.empty(m, dtype=dt).numpy())
            if name in ("count", "coverage"):
                ho = alloc(n, torch.int64)
                hfn = ix.count_overlaps if name == "count" else ix.coverage
                return (lambda: hfn(k_, s_, e_, out=ho)), (lambda r: (ho,)), n * 12, n * 8
            if name == "nearest":
                o3 = (alloc(n, torch.int32).view(np.uint32), alloc(((n + 31) // 32) * 4, torch.uint8), alloc(n, torch.int64))
                return (lambda: ix.nearest_k1_into(k_, s_, e_, o3)), (lambda r: o3), n * 12, n * 12 + n // 8
            # pairs: results stay in the library's pinned buffers (zero-copy views)
            return (lambda: ix.overlap_pairs(k_, s_, e_, mode=brg.PAIRS_INNER_ANY_ORDER, copy=False)), (lambda r: r), n * 12, units_per_step * 8

        def time_e2e(estep):
            # warm-up: two results alive at once, as in the timed loop (`r` still holds step k - 1 while step k runs), so that
            # the library's pinned result pool and staging ring hold everything before the clock starts
            w0 = estep()
            w1 = estep()
            del w0, w1
            h.barrier()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            t0 = time.perf_counter()
            e0.record()
            for _ in range(e2e_steps):
                r = estep()
            e1.record()
            torch.cuda.synchronize()
            wall_ms = (time.perf_counter() - t0) * 1e3
            return h.reduce(max(e0.elapsed_time(e1), wall_ms), "max") / e2e_steps, r

        estep, result_of, h2d, d2h = make_step(npk, nps, npe, True)
        e2e_ms, r = time_e2e(estep)
        e_units = units_per_step
        res["e2e"] = {"value": h.reduce(e_units, "sum") / (e2e_ms * 1e-3), "unit": res["unit"], "h2d_bytes_per_step": h2d,
                      "d2h_bytes_per_step": d2h, "ms_per_step": e2e_ms,
                      "api": API[name] + " (pinned host buffers; per-row key id column in, the library scans it for runs chunk by chunk"
                                         " inside the timed region and sends runs where a chunk has <= 1024 of them)",
                      "h2d_bytes_note": "counted as the caller's columns (12 B/row); with key runs 8 B/row cross PCIe"}
        if not args.no_parity:
            got = result_of(r)
            ok = check_pairs(h, oi, pnp, got[0], got[1]) if name == "pairs" else check_fixed(h, name, oi, pnp, got)
            parity["e2e"] = bool(h.reduce(1.0 if ok else 0.0, "min") == 1.0)
        del r
        if want_pageable:
            # ordinary heap memory, as arrow-rs allocates it: the library stages it through its pinned ring with host threads
            gk, gs, ge = npk.copy(), nps.copy(), npe.copy()
            estep2, result_of2, _, _ = make_step(gk, gs, ge, False)
            pg_ms, r2 = time_e2e(estep2)
            res["e2e_pageable"] = {"value": h.reduce(e_units, "sum") / (pg_ms * 1e-3), "unit": res["unit"], "ms_per_step": pg_ms,
                                   "what": "same call, pageable (heap) input and output buffers: copied through the library's pinned "
                                           "staging ring by its host threads (BRG_HOST_THREADS)"}
            if not args.no_parity:
                got = result_of2(r2)
                ok = check_pairs(h, oi, pnp, got[0], got[1]) if name == "pairs" else check_fixed(h, name, oi, pnp, got)
                parity["e2e_pageable"] = bool(h.reduce(1.0 if ok else 0.0, "min") == 1.0)
            del r2, gk, gs, ge
        res["_host_cols"] = (npk, nps, npe)
    if parity:
        res["parity"] = parity
        res["parity_all"] = all(parity.values())
        if not res["parity_all"]:
            raise SystemExit(f"bench result of {name}:{order} differs from the oracle ({parity}): refusing to report a number")
    res["_build_cols"] = b
    ix.close()
    return res


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="count", choices=list(WORKLOADS))
    ap.add_argument("--order", default=os.environ.get("BRG_BENCH_ORDER", "sorted"), choices=["sorted", "shuffled"])
    ap.add_argument("--scale", type=float, default=float(os.environ.get("BRG_BENCH_SCALE", "1.0")),
                    help="shrink the workload (debug only; the judged number uses 1.0)")
    ap.add_argument("--cpu-sample", type=int, default=32_000_000)
    ap.add_argument("--no-extras", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-parity", action="store_true", help="skip the oracle comparison (profiling runs only)")
    ap.add_argument("--no-e2e", action="store_true", help="device-resident steps only (profiling runs only)")
    ap.add_argument("--affinity", default=os.environ.get("BRG_BENCH_AFFINITY", "split"), choices=["split", "nvml", "none"])
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup

    import torch
    import torch.distributed as dist

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    name = args.workload
    metric = "overlap pairs/sec" if name == "pairs" else "probe intervals/sec"
    unit = "pairs/s" if name == "pairs" else "probes/s"
    config = {"workload": WORKLOADS[name][2], "probe_order": args.order, "probes_per_gpu": int(WORKLOADS[name][0] * args.scale),
              "build_rows": int(WORKLOADS[name][1] * args.scale), "contigs": len(HG38),
              "l2": "inputs larger than L2 (>=1.5 GB streamed per step)", "parallelism": f"probe-row shards x{world}, build replicated"}

    # ------------------------------------------------------------------------------------------------------
    if args.impl == "reference":
        if rank != 0:
            return
        dev = "cuda:0" if torch.cuda.is_available() else "cpu"
        b, p = make_workload(torch, name, dev, 0, max(world, 1), args.order, args.scale)
        sample = min(args.cpu_sample, p[0].numel())
        bnp = tuple(to_np(x).astype(np.uint32) if i == 0 else to_np(x) for i, x in enumerate(b))
        pnp = tuple(to_np(x[:sample]).astype(np.uint32) if i == 0 else to_np(x[:sample]) for i, x in enumerate(p))
        v, cores, desc, dt = cpu_run(name, bnp, pnp, sample, max(args.steps, 1), args.warmup)
        line = {"impl": "reference", "metric": metric, "value": v, "unit": unit, "n_gpus": args.gpus, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "int32", "data": "synthetic", "config": config,
                "cpu_baseline": {"value": v, "unit": unit, "cores": cores, "kind": "port", "sample": desc},
                "e2e": {"value": v, "unit": unit, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        print(json.dumps(line))
        return

    # ------------------------------------------------------------------------------------------------------
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback in the product path)")
    torch.cuda.set_device(local_rank)
    device = f"cuda:{local_rank}"
    if world > 1 and args.affinity != "none":  # (N = 1 keeps every host core for the CPU arm)
        if args.affinity == "nvml":
            from bio_ranges_b200.sharding import bind_near_gpu

            cores = bind_near_gpu(local_rank)
            config["host_affinity"] = f"each rank bound to the {cores} cores NVML names for its GPU" if cores else "unbound"
        else:
            cores = split_affinity(local_rank, world)
            config["host_affinity"] = f"each rank bound to its own {cores} of the job's cores" if cores else "unbound"
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device(device))
    import bio_ranges_b200 as brg

    h = Harness(torch, dist, brg, rank, local_rank, world, device, args)
    sampler = ClockSampler(local_rank)
    head = run_workload(h, name, args.order, args.steps, args.warmup, want_e2e=not args.no_e2e, want_pageable=True, want_build_fig=True,
                        sampler=sampler)
    # the sampler (20 ms period) ran through both timed regions: the device-resident steps and the end-to-end steps
    clocks = sampler.stop() if rank == 0 else None

    # ---- CPU baseline on a bounded sample (rank 0, N=1) ----------------------------------------------------------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu and "_host_cols" in head:
        npk, nps, npe = head["_host_cols"]
        sample = min(args.cpu_sample, len(nps))
        b = head["_build_cols"]
        bnp = (to_np(b[0]).astype(np.uint32), to_np(b[1]), to_np(b[2]))
        pnp = (npk[:sample], nps[:sample], npe[:sample])
        v, cores, desc, _ = cpu_run(name, bnp, pnp, sample, 2, 1)
        cpu = {"value": v, "unit": unit, "cores": cores, "kind": "port", "sample": desc}
    head.pop("_host_cols", None)
    head.pop("_build_cols", None)
    h._oracle_cache.clear()

    # ---- the other section-8 workloads: every rank runs its shard, each with its own parity (reported as extras) -----------
    extras = {}
    if not args.no_extras and name == "count":
        alt = "shuffled" if args.order == "sorted" else "sorted"
        plan = [("count", alt, False), ("coverage", args.order, True), ("nearest", args.order, True), ("pairs", args.order, True)]
        if world == 1:
            plan += [("coverage", alt, False), ("nearest", alt, False), ("pairs", alt, False)]
        for wname, worder, we2e in plan:
            key = wname if worder == args.order else f"{wname}_{worder}"
            try:
                r = run_workload(h, wname, worder, 5, 3, want_e2e=we2e and not args.no_e2e)
                r.pop("_host_cols", None)
                r.pop("_build_cols", None)
                r["roofline_frac"] = r["roofline"]["frac"]
                extras[key] = r
            except SystemExit:
                raise
            except Exception as ex:  # an extra that cannot run never invalidates the headline
                extras[key] = {"error": repr(ex)}
            h._oracle_cache.clear()
            torch.cuda.empty_cache()
        if world == 1 and rank == 0:
            try:
                extras["merge"] = quick_merge_run(torch, h.ctx, args.scale, device)
                extras["nearest_k5"] = quick_nearest_k_run(torch, h, 5)
            except Exception as ex:
                extras["error"] = repr(ex)

    if rank == 0:
        parity_all = bool(head.get("parity_all", False)) and all(v.get("parity_all", True) for v in extras.values() if isinstance(v, dict))
        line = {"metric": metric, "value": head["value"], "unit": unit, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": head["ms_per_step"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "int32",
                "data": "synthetic", "config": config, "clocks": clocks, "e2e": head.get("e2e"), "e2e_pageable": head.get("e2e_pageable"),
                "gpu_launches": head["gpu_launches"], "roofline": head["roofline"], "roofline_with_build": head.get("roofline_with_build"),
                "cpu_baseline": cpu, "build": head["build"],
                "parity": None if args.no_parity else {"all_ranks_all_rows": parity_all, "headline": head.get("parity"),
                                                       "how": "every rank: all rows of the device-resident output and of the end-to-end "
                                                              "output vs the oracle; min over ranks"},
                "parity_vs_oracle_sample": None if args.no_parity else parity_all, "extras": extras}
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def quick_merge_run(torch, ctx, scale, device):
    """merge UDTF core through the host-buffer entry point brg_merge: 10M intervals, 24 contigs"""
    n = max(1024, int(10_000_000 * scale))
    k, s, e = gen_intervals(torch, n, 44, 1, 1550, device)
    hk, hs, he = (to_np(k).astype(np.uint32), to_np(s), to_np(e))
    ctx.merge(hk, hs, he, len(HG38))
    t0 = time.perf_counter()
    reps = 3
    for _ in range(reps):
        out = ctx.merge(hk, hs, he, len(HG38))
    dt = (time.perf_counter() - t0) / reps
    return {"workload": f"merge {n} intervals, 24 contigs, min_dist 0 (host buffers -> sort -> running max -> compact -> host)",
            "value": n / dt, "unit": "intervals/s", "ms_per_step": dt * 1e3, "merged_rows": int(len(out[0]))}


def quick_nearest_k_run(torch, h, k):
    """nearest with k > 1 (nearest_index.rs:103-190) through brg_nearest: 5M queries x 5M sparse targets, host buffers"""
    b, p = make_workload(torch, "nearest", h.device, 0, 1, "sorted", h.args.scale * 0.1)
    b = gen_intervals(torch, max(1024, int(5_000_000 * h.args.scale)), 42, 50, 500, h.device)
    ix = h.ctx.build_dev(*b, len(HG38), h.brg.WANT_NEAREST)
    npk, nps, npe = to_np(p[0]).astype(np.uint32), to_np(p[1]), to_np(p[2])
    ix.nearest_k(npk, nps, npe, k)
    t0 = time.perf_counter()
    reps = 3
    for _ in range(reps):
        out = ix.nearest_k(npk, nps, npe, k)
    dt = (time.perf_counter() - t0) / reps
    ix.close()
    return {"workload": f"nearest k={k}: {len(nps)} sorted queries x {b[1].numel()} sparse targets (host buffers, count -> scan -> emit, "
                        "one row per thread)", "value": len(nps) / dt, "unit": "probes/s", "ms_per_step": dt * 1e3, "rows_out": int(len(out[0]))}


if __name__ == "__main__":
    main()
