"""Bins tables on SHUFFLED probe tables at bench.py's full sizes, device-resident (CUDA events, no profiler):
  count    125M x 10M  : count bins wide (BRG_CB_PACK=0) vs packed (=1)
  nearest  50M  x 5M   : rank directories (BRG_NB_BINS=0) vs nearest / pair bins (=1)
  pairs    10M  x 10M  : three-launch chunked form, range pass through the directories vs the bins (BRG_NB_BINS=2), wide vs packed
Every variant: warm build time, index bytes, the bins tables' shapes, ms per step, and a parity sample against the oracle.
(the switches are read at every build, so one process runs all variants)"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch

import bench
import bio_ranges_b200 as brg
from oracle import oracle as O

PEAK = bench.hbm_peak()[0]
ctx = brg.Context(0)
dev = "cuda:0"
SAMPLE = 7813 * 256  # whole 256-row chunks (the pair offsets are per chunk)
only = sys.argv[1:] or ["count", "nearest", "pairs"]


def timed(fn, reps=5):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def build(b, want, env):
    for k in ("BRG_CB_PACK", "BRG_NB_BINS"):
        os.environ.pop(k, None)
    os.environ.update(env)
    ix = ctx.build_dev(*b, len(bench.HG38), want)
    torch.cuda.synchronize()
    ix.close()

    def again():
        ctx.build_dev(*b, len(bench.HG38), want).close()

    ms = timed(again, 3)
    ix = ctx.build_dev(*b, len(bench.HG38), want)
    torch.cuda.synchronize()
    return ix, ms


def report(name, env, ix, build_ms, ms, alg_bytes, ok, extra=None):
    r = {"workload": name, "env": env, "build_ms_warm": round(build_ms, 3), "index_mb": round(ix.device_bytes / 1e6, 1),
         "bins": {t: ix.bins_info(i) for i, t in enumerate(("count", "coverage", "nearest_pairs")) if ix.bins_info(i)["records"]},
         "ms": round(ms, 4), "roofline_frac": round(alg_bytes / (ms * 1e-3) / 1e9 / PEAK, 4), "parity_sample": ok}
    if extra:
        r.update(extra)
    print(json.dumps(r), flush=True)


def host(t, m=None):
    return t[:m].cpu().numpy() if m else t.cpu().numpy()


for name in only:
    b, p = bench.make_workload(torch, name, dev, 0, 1, "shuffled", 1.0)
    pk, ps, pe = p
    n, nb = ps.numel(), b[1].numel()
    oi = O.OracleIndex(host(b[0]).astype(np.uint32), host(b[1]), host(b[2]), len(bench.HG38))
    hk, hs, he = host(pk, SAMPLE).astype(np.uint32), host(ps, SAMPLE), host(pe, SAMPLE)
    thr = O.max_threads()
    if name == "count":
        want = oi.count_overlaps(hk, hs, he, threads=thr)
        out = torch.empty(n, dtype=torch.int64, device=dev)
        for env in ({"BRG_CB_PACK": "0"}, {"BRG_CB_PACK": "1"}, {}):
            ix, bms = build(b, brg.WANT_COUNT, env)
            ms = timed(lambda: ix.count_overlaps_dev(pk, ps, pe, out))
            ok = bool((host(out, SAMPLE) == want).all())
            report(name, env, ix, bms, ms, n * 20 + nb * 8, ok)
            ix.close()
        del out
    elif name == "nearest":
        wl, wv, wd = oi.nearest_k1(hk, hs, he, threads=thr)
        left = torch.empty(n, dtype=torch.int32, device=dev)
        bits = torch.zeros((n + 31) // 32, dtype=torch.int32, device=dev)
        dd = torch.empty(n, dtype=torch.int64, device=dev)
        for env in ({"BRG_NB_BINS": "0"}, {"BRG_NB_BINS": "1"}, {"BRG_NB_BINS": "1", "BRG_CB_PACK": "0"}):
            ix, bms = build(b, brg.WANT_NEAREST, env)
            ms = timed(lambda: ix.nearest_k1_dev(pk, ps, pe, left, bits, dd))
            gv = np.unpackbits(host(bits, SAMPLE // 32).view(np.uint8), bitorder="little").astype(bool)
            gl, gd = host(left, SAMPLE).view(np.uint32), host(dd, SAMPLE)
            ok = bool((gv == wv).all() and (gl[wv] == wl[wv]).all() and (gd[wv] == wd[wv]).all())
            report(name, env, ix, bms, ms, n * 28 + n // 8 + nb * 12, ok)
            ix.close()
        del left, bits, dd
    else:
        cnt = oi.overlap_counts(hk, hs, he, threads=thr)
        offsets = torch.empty((n + 255) // 256 + 1, dtype=torch.int64, device=dev)
        ranges = torch.empty(n, dtype=torch.int64, device=dev)
        for env in ({"BRG_NB_BINS": "0"}, {"BRG_NB_BINS": "2", "BRG_CB_PACK": "0"}, {"BRG_NB_BINS": "2"}):  # (2: also a pairs-only index)
            ix, bms = build(b, brg.WANT_PAIRS, env)
            ix.pairs_chunk_offsets_dev(pk, ps, pe, offsets, ranges)
            torch.cuda.synchronize()
            total = int(offsets[-1].item())
            bi = torch.empty(max(total, 1), dtype=torch.int32, device=dev)
            pi = torch.empty(max(total, 1), dtype=torch.int32, device=dev)
            ms1 = timed(lambda: ix.pairs_chunk_offsets_dev(pk, ps, pe, offsets, ranges))
            ms = timed(lambda: (ix.pairs_chunk_offsets_dev(pk, ps, pe, offsets, ranges), ix.pairs_fill_dev(ps, offsets, ranges, bi, pi)))
            torch.cuda.synchronize()
            # parity sample: pairs of the first SAMPLE probe rows are the first offsets[SAMPLE / 256] output slots (probe-major)
            m = int(offsets[SAMPLE // 256].item())
            gp = host(pi, m).view(np.uint32)
            ok = bool(m == int(cnt.sum()) and (np.bincount(gp, minlength=SAMPLE)[:SAMPLE] == cnt).all())
            report(name, env, ix, bms, ms, n * 12 + total * 8 + nb * 12, ok, {"range_pass_ms": round(ms1, 4), "pairs": total})
            del bi, pi
            ix.close()
        del offsets, ranges
    del b, p, pk, ps, pe, oi
    torch.cuda.empty_cache()
