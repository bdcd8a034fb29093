#!/usr/bin/env python
"""bench.py -- throughput of the interval-overlap hot path on N B200s (one process per GPU).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--workload count|pairs|coverage|nearest] [--order sorted|shuffled]
  python bench.py --impl reference ...      # the reference's CPU algorithm (oracle port) on the host cores

Headline workload (BASELINE.json north_star / configs[4]): count_overlaps, 125M probe intervals per GPU x 10M build
intervals over 24 hg38-like contigs (N=8 => 1B x 10M), weak scaling: the build index is built on rank 0 and replicated
with an NCCL broadcast, probe rows are sharded by contiguous row range, no collective on the data path.

One JSON line on rank 0:
  value     probes/s, whole job, inputs resident in HBM, K timed steps bracketed by barrier + synchronize, max over ranks
  e2e       same metric through the host-buffer C-ABI call (brg_count_overlaps) with pinned HOST buffers: H2D + kernels +
            D2H inside the timed region
  roofline  dominant kernel vs the measured HBM copy bandwidth (MEASURED_PEAKS.json)
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


def hbm_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def ncu_traffic(kernel, workload, args):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of the dominant kernel, from the committed
    `ncu --set full` capture of this same workload (profiles/traffic.json); None when the run differs from it."""
    p = os.path.join(ROOT, "profiles", "traffic.json")
    if not os.path.exists(p) or args.scale != 1.0:
        return None
    rec = json.load(open(p)).get(f"{workload}:{args.order}")
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
        fn = lambda: oi.overlap_counts(pk, ps, pe, threads=cores)  # tree walk visiting every pair (no output vector)
    for _ in range(warmup):
        fn()
    t0 = time.perf_counter()
    for _ in range(steps):
        res = fn()
    dt = (time.perf_counter() - t0) / steps
    units = len(ps)
    if name == "pairs":
        units = int(res.sum())
    return units / dt, cores, f"first {len(ps)} probe rows of the workload, {steps} timed passes, index build {t_build:.2f}s excluded", dt


def to_np(t):
    return t.cpu().numpy()


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
    near_cores = 0
    if world > 1:  # (N = 1 keeps every host core for the CPU arm)
        from bio_ranges_b200.sharding import bind_near_gpu

        near_cores = bind_near_gpu(local_rank)
        config["host_affinity"] = f"each rank bound to the {near_cores} cores next to its GPU (NVML)" if near_cores else "unbound"
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device(device))
    import bio_ranges_b200 as brg

    ctx = brg.Context(local_rank)
    lib = brg._cabi.lib()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(ms):
        if world == 1:
            return ms
        t = torch.tensor([ms], dtype=torch.float64, device=device)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def sum_over_ranks(v):
        if world == 1:
            return v
        t = torch.tensor([v], dtype=torch.float64, device=device)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

    b, p = make_workload(torch, name, device, rank, world, args.order, args.scale)
    want = {"count": brg.WANT_COUNT, "pairs": brg.WANT_PAIRS, "coverage": brg.WANT_COVERAGE, "nearest": brg.WANT_NEAREST}[name]

    # ---- build on rank 0, replicate with an NCCL broadcast of the index slab --------------------------------
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    if rank == 0:
        ix = ctx.build_dev(*b, len(HG38), want)
        torch.cuda.synchronize()
    build_s = time.perf_counter() - t0
    bcast_ms = 0.0
    if world > 1:
        from bio_ranges_b200.sharding import replicate_index

        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        ix = replicate_index(ctx, ix if rank == 0 else None, dist, rank, world, device)  # header + one slab broadcast
        e1.record()
        torch.cuda.synchronize()
        bcast_ms = max_over_ranks(e0.elapsed_time(e1))
    pk, ps, pe = p
    n = ps.numel()

    # ---- device-resident step ------------------------------------------------------------------------------
    if name in ("count", "coverage"):
        out = torch.empty(n, dtype=torch.int64, device=device)
        fn = (ix.count_overlaps_dev if name == "count" else ix.coverage_dev)
        step = lambda: fn(pk, ps, pe, out)
        units_per_step = n
        alg_bytes = n * 20 + int(WORKLOADS[name][1] * args.scale) * 8
        launches_per_step = 1
    elif name == "nearest":
        left = torch.empty(n, dtype=torch.int32, device=device)
        bits = torch.empty((n + 31) // 32, dtype=torch.int32, device=device)
        dd = torch.empty(n, dtype=torch.int64, device=device)
        step = lambda: ix.nearest_k1_dev(pk, ps, pe, left, bits, dd)
        units_per_step = n
        alg_bytes = n * 28 + n // 8 + int(WORKLOADS[name][1] * args.scale) * 12
        launches_per_step = 1
    else:
        offsets = torch.empty((n + 255) // 256 + 1, dtype=torch.int64, device=device)  # per 256-row chunk
        ranges = torch.empty(n, dtype=torch.int64, device=device)  # per-row candidate ranges (pass 1 -> pass 2)
        ix.pairs_chunk_offsets_dev(pk, ps, pe, offsets, ranges)
        torch.cuda.synchronize()
        total = int(offsets[-1].item())
        bi = torch.empty(max(total, 1), dtype=torch.int32, device=device)
        pi = torch.empty(max(total, 1), dtype=torch.int32, device=device)

        def step():  # count -> scan -> emit
            ix.pairs_chunk_offsets_dev(pk, ps, pe, offsets, ranges)
            ix.pairs_fill_dev(ps, offsets, ranges, bi, pi)

        units_per_step = total
        alg_bytes = n * 12 + total * 8 + int(WORKLOADS[name][1] * args.scale) * 12
        launches_per_step = 7
    for _ in range(args.warmup):
        step()
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    l0 = lib.brg_kernel_launches()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps + 1)]
    barrier()
    ev[0].record()
    for i in range(args.steps):
        step()
        ev[i + 1].record()
    torch.cuda.synchronize()
    barrier()
    total_ms = max_over_ranks(ev[0].elapsed_time(ev[-1]))
    launches = lib.brg_kernel_launches() - l0
    per_step_ms = [ev[i].elapsed_time(ev[i + 1]) for i in range(args.steps)]
    ms_per_step = total_ms / args.steps
    value = sum_over_ranks(units_per_step) / (ms_per_step * 1e-3)
    peak, peak_src = hbm_peak()
    kern_ms = float(np.mean(per_step_ms))
    achieved = alg_bytes / (kern_ms * 1e-3) / 1e9
    kernel_name = {"count": "count_kernel_tiled", "coverage": "coverage_kernel_tiled", "nearest": "nearest_k1_kernel_tiled",
                   "pairs": "pairs_ranges_kernel_tiled+scan+pairs_emit_kernel_flat"}[name]
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": ncu_traffic(kernel_name, name, args), "peak_source": peak_src, "kernel": kernel_name,
                "algorithmic_bytes_per_launch": alg_bytes, "launch_ms": kern_ms}

    # ---- end to end through the host-buffer C-ABI call: pinned host inputs -> H2D -> kernels -> D2H ----------
    e2e = None
    h_pk = torch.empty(n, dtype=torch.int32).pin_memory(); h_pk.copy_(pk)
    h_ps = torch.empty(n, dtype=torch.int32).pin_memory(); h_ps.copy_(ps)
    h_pe = torch.empty(n, dtype=torch.int32).pin_memory(); h_pe.copy_(pe)
    npk, nps, npe = h_pk.numpy().view(np.uint32), h_ps.numpy(), h_pe.numpy()
    e2e_steps = max(3, min(args.steps, 5))
    # The contig column travels in run-length form (brg_*_runs): the host's contig-cache loop (interval_tree.rs:258-265)
    # yields runs on a coordinate-sorted table; a shuffled table has no runs to speak of and sends the id column.
    run_key, run_end = brg._cabi.key_runs(npk)
    use_runs = len(run_key) <= max(1024, n // 4096)
    if name in ("count", "coverage"):
        h_out = torch.empty(n, dtype=torch.int64).pin_memory()
        if use_runs:
            hfn = ix.count_overlaps_runs if name == "count" else ix.coverage_runs
            estep = lambda: hfn(run_key, run_end, nps, npe, out=h_out.numpy())
        else:
            hfn = ix.count_overlaps if name == "count" else ix.coverage
            estep = lambda: hfn(npk, nps, npe, out=h_out.numpy())
        h2d, d2h, e_units = n * (8 if use_runs else 12), n * 8, n
    elif name == "nearest":
        if not use_runs:  # the id column form has no caller-buffer variant in the python wrapper: send one run per row change
            use_runs = True
        h_left = torch.empty(n, dtype=torch.int32).pin_memory()
        h_bits = torch.zeros(((n + 31) // 32) * 4, dtype=torch.uint8).pin_memory()
        h_dist = torch.empty(n, dtype=torch.int64).pin_memory()
        nout = (h_left.numpy().view(np.uint32), h_bits.numpy(), h_dist.numpy())
        estep = lambda: ix.nearest_k1_runs(run_key, run_end, nps, npe, out=nout)
        h2d, d2h, e_units = n * (8 if use_runs else 12), n * 12 + n // 8, n
    else:
        use_runs = False
        estep = lambda: ix.overlap_pairs(npk, nps, npe, copy=False)  # results stay in the library's pinned buffers (zero-copy views)
        h2d, d2h, e_units = n * 12, units_per_step * 8, units_per_step
    # warm-up: two results alive at once, as in the timed loop below (`r` still holds step k - 1 while step k runs), so that
    # the library's pinned result pool holds both sets before the clock starts
    w0 = estep()
    w1 = estep()
    del w0, w1
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record()
    for _ in range(e2e_steps):
        r = estep()
    e1.record()
    torch.cuda.synchronize()
    wall_ms = (time.perf_counter() - t0) * 1e3
    e2e_ms = max_over_ranks(max(e0.elapsed_time(e1), wall_ms)) / e2e_steps
    # the sampler (100 ms period) ran through both timed regions: the device-resident steps and the end-to-end steps
    clocks = sampler.stop() if rank == 0 else None
    e2e = {"value": sum_over_ranks(e_units) / (e2e_ms * 1e-3), "unit": unit, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
           "ms_per_step": e2e_ms, "api": {"count": "brg_count_overlaps", "coverage": "brg_coverage", "nearest": "brg_nearest_k1",
                                          "pairs": "brg_overlap_pairs"}[name] + ("_runs (host buffers, pinned; contig column as key runs)"
                                                                                if use_runs else " (host buffers, pinned)")}
    # parity spot check of the timed result against the oracle (rank 0, small sample) -- the checker, not the product
    parity = None
    if rank == 0 and name in ("count", "coverage") and not args.no_cpu:
        from oracle import oracle as O

        m = min(200_000, n)
        sel = slice(0, m)
        oi = O.OracleIndex(to_np(b[0]).astype(np.uint32), to_np(b[1]), to_np(b[2]), len(HG38))
        got = h_out.numpy()[sel]
        wantv = (oi.count_overlaps if name == "count" else oi.coverage)(npk[sel], nps[sel], npe[sel], threads=O.max_threads())
        parity = bool((got == wantv).all())
        if not parity:
            raise SystemExit("bench result differs from the oracle: refusing to report a number")

    # ---- CPU baseline on a bounded sample (rank 0, N=1) ----------------------------------------------------------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        sample = min(args.cpu_sample, n)
        bnp = (to_np(b[0]).astype(np.uint32), to_np(b[1]), to_np(b[2]))
        pnp = (npk[:sample], nps[:sample], npe[:sample])
        v, cores, desc, _ = cpu_run(name, bnp, pnp, sample, 2, 1)
        cpu = {"value": v, "unit": unit, "cores": cores, "kind": "port", "sample": desc}

    # ---- other section-8 workloads, device-resident, short (N=1 only; reported as extras) ------------------------
    extras = {}
    if rank == 0 and world == 1 and not args.no_extras and name == "count":
        del h_pk, h_ps, h_pe, h_out
        alt = "shuffled" if args.order == "sorted" else "sorted"
        try:
            extras[f"count_{alt}"] = quick_device_run(torch, brg, ctx, "count", alt, args.scale, device, peak, 5)
            for other in ("coverage", "nearest", "pairs"):
                extras[other] = quick_device_run(torch, brg, ctx, other, args.order, args.scale, device, peak, 5)
            extras["merge"] = quick_merge_run(torch, ctx, args.scale, device)
        except Exception as ex:  # extras never invalidate the headline
            extras["error"] = repr(ex)

    if rank == 0:
        line = {"metric": metric, "value": value, "unit": unit, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "int32",
                "data": "synthetic", "config": config, "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches),
                "roofline": roofline, "cpu_baseline": cpu, "build": {"seconds_rank0": build_s, "index_bytes": ix.device_bytes,
                                                                      "broadcast_ms": bcast_ms},
                "parity_vs_oracle_sample": parity, "extras": extras}
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


def quick_device_run(torch, brg, ctx, name, order, scale, device, peak, steps):
    b, p = make_workload(torch, name, device, 0, 1, order, scale)
    want = {"count": brg.WANT_COUNT, "pairs": brg.WANT_PAIRS, "coverage": brg.WANT_COVERAGE, "nearest": brg.WANT_NEAREST}[name]
    ix = ctx.build_dev(*b, len(HG38), want)
    pk, ps, pe = p
    n = ps.numel()
    nb = b[1].numel()
    if name in ("count", "coverage"):
        out = torch.empty(n, dtype=torch.int64, device=device)
        fn = ix.count_overlaps_dev if name == "count" else ix.coverage_dev
        step = lambda: fn(pk, ps, pe, out)
        units, alg = n, n * 20 + nb * 8
    elif name == "nearest":
        left = torch.empty(n, dtype=torch.int32, device=device)
        bits = torch.empty((n + 31) // 32, dtype=torch.int32, device=device)
        dd = torch.empty(n, dtype=torch.int64, device=device)
        step = lambda: ix.nearest_k1_dev(pk, ps, pe, left, bits, dd)
        units, alg = n, n * 28 + n // 8 + nb * 12
    else:
        offsets = torch.empty((n + 255) // 256 + 1, dtype=torch.int64, device=device)  # per 256-row chunk
        ranges = torch.empty(n, dtype=torch.int64, device=device)  # per-row candidate ranges (pass 1 -> pass 2)
        ix.pairs_chunk_offsets_dev(pk, ps, pe, offsets, ranges)
        torch.cuda.synchronize()
        total = int(offsets[-1].item())
        bi = torch.empty(max(total, 1), dtype=torch.int32, device=device)
        pi = torch.empty(max(total, 1), dtype=torch.int32, device=device)

        def step():  # count -> scan -> emit
            ix.pairs_chunk_offsets_dev(pk, ps, pe, offsets, ranges)
            ix.pairs_fill_dev(ps, offsets, ranges, bi, pi)

        units, alg = total, n * 12 + total * 8 + nb * 12
    for _ in range(3):
        step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        step()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    ix.close()
    return {"workload": WORKLOADS[name][2], "probe_order": order, "value": units / (ms * 1e-3), "unit": "pairs/s" if name == "pairs" else "probes/s",
            "ms_per_step": ms, "roofline_frac": alg / (ms * 1e-3) / 1e9 / peak, "algorithmic_bytes": alg,
            **({"pairs_per_probe": units / n} if name == "pairs" else {})}


if __name__ == "__main__":
    main()
