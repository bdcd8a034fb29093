"""stage times of the index build (BRG_BUILD_TRACE=1: every stage synchronised, wall clock) with and without the nearest / pair
bins, on bench.py's build sides.  usage: BRG_BUILD_TRACE=1 python profiles/exp_build_trace.py [pairs|nearest|count ...]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch

import bench
import bio_ranges_b200 as brg

ctx = brg.Context(0)
for name in sys.argv[1:] or ["pairs"]:
    b, _ = bench.make_workload(torch, name, "cuda:0", 0, 1, "sorted", 1.0)
    want = {"count": brg.WANT_COUNT, "pairs": brg.WANT_PAIRS, "coverage": brg.WANT_COVERAGE, "nearest": brg.WANT_NEAREST}[name]
    for nb in ("0", "1"):
        os.environ["BRG_NB_BINS"] = nb
        for rep in range(3):
            sys.stderr.write(f"---- {name} build, BRG_NB_BINS={nb}, call {rep}\n")
            sys.stderr.flush()
            ix = ctx.build_dev(*b, len(bench.HG38), want)
            torch.cuda.synchronize()
            ix.close()
