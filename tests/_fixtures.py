"""Literal copies of the reference's tiny TEST DATA tables (not source code):
testing/data/interval/{reads,targets}.csv and testing/data/ranges/{reads,targets}.csv, plus helpers that
dictionary-encode contig strings to dense uint32 key ids the way the host layer does."""
import os

import numpy as np

ABSENT = 0xFFFFFFFF

# testing/data/ranges/reads.csv (13 rows; interval/reads.csv is the same without the chr3 row)
RANGES_READS = [
    ("chr1", 150, 250), ("chr1", 190, 300), ("chr1", 300, 501), ("chr1", 500, 700), ("chr1", 22000, 22300),
    ("chr1", 15000, 15000), ("chr2", 150, 250), ("chr2", 190, 300), ("chr2", 300, 500), ("chr2", 500, 700),
    ("chr2", 22000, 22300), ("chr2", 15000, 15000), ("chr3", 234, 300),
]
# testing/data/ranges/targets.csv (11 rows; interval/targets.csv is the same without the chr3 row)
RANGES_TARGETS = [
    ("chr1", 100, 190), ("chr1", 200, 290), ("chr1", 400, 600), ("chr1", 10000, 20000), ("chr1", 22100, 22100),
    ("chr2", 100, 190), ("chr2", 200, 290), ("chr2", 400, 600), ("chr2", 10000, 20000), ("chr2", 22100, 22100),
    ("chr3", 100, 200),
]
INTERVAL_READS = RANGES_READS[:12]
INTERVAL_TARGETS = RANGES_TARGETS[:10]


def columns(rows):
    contig = [r[0] for r in rows]
    start = np.array([r[-2] for r in rows], np.int32)
    end = np.array([r[-1] for r in rows], np.int32)
    return contig, start, end


def encode(build_contigs, probe_contigs):
    """Dense ids over the BUILD side's distinct keys; probe keys absent from the build map to ABSENT
    (the reference's `indexes.get(contig)` returning None: IT:265, NE:354, IJ:857)."""
    d = {}
    for c in build_contigs:
        d.setdefault(c, len(d))
    bk = np.array([d[c] for c in build_contigs], np.uint32)
    pk = np.array([d.get(c, ABSENT) for c in probe_contigs], np.uint32)
    return bk, pk, len(d)


def golden():
    p = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ranges_golden.npz")
    return np.load(p)


def rows_set(*cols):
    return sorted(zip(*[list(map(lambda v: v.item() if hasattr(v, "item") else v, c)) for c in cols]))
