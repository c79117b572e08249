"""Generate tests/golden/encoder_cpp.npz with the reference's own COMPILED C++ feature encoder.

Run in the build container only (needs /root/reference to build oracle/_ref/libref_encoder.so):

    make -C oracle && PYTHONDONTWRITEBYTECODE=1 python oracle/gen_encoder_cpp_golden.py

oracle/ref_encoder.py calls ``AlleleSearcherLiteFiltered::computeFeaturesColoredSimple``
(/root/reference/c++/src/AlleleSearcherLiteFiltered.cpp:1031-1180) through oracle/ref_encoder_driver.cpp.  Unlike the
reference's Python specification (tests/golden/encoder.npz) the C++ covers every input, so the cases here are the
ones the specification cannot pin: soft and hard clips, insertions / deletions at the window borders and at the start
of a read, insertions into reads of varying quality, N bases, reads hanging over both window ends, the technology
filter and the no-support dummy row.  Stored: whole sites (all reads + the support lists) and the C++ output per
(site, allele, technology, hp) query.
"""
from __future__ import annotations

import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "tests", "golden", "encoder_cpp.npz")
sys.path.insert(0, ROOT)


def handmade_sites(E):
    """Corner cases by construction.  Window of a site = [mid - 75, mid + 75), mid = (a0 + a1) // 2."""
    ref = "".join("ACGT"[(i * 7 + i // 3) % 4] for i in range(600))
    S = E.SitePileup
    M, I, D, N, SC, HC, P, EQ, X = range(9)
    sites = []
    a0, a1 = 300, 303                      # window [226, 376)
    reads = [
        ("ACGTACGTACGT", [(I, 3), (M, 9)], 230),                 # read opens with an insertion (no base before it)
        ("ACGTACGTACGT", [(D, 2), (M, 12)], 230),                # read opens with a deletion (quality of base -1)
        ("ACGTACGTACGT", [(SC, 4), (M, 8)], 226),                # soft clip, first aligned base on the left border
        ("ACGTACGTACGT", [(HC, 5), (M, 12)], 370),               # hard clip; read runs over the right border
        ("ACGTACGTACGT", [(M, 4), (D, 5), (M, 8)], 222),         # deletion straddles the left border (base before it outside)
        ("ACGTACGTACGT", [(M, 5), (D, 5), (M, 7)], 222),         # base before the deletion is the first window column
        ("ACGTACGTACGT", [(M, 6), (D, 9), (M, 6)], 366),         # deletion straddles the right border
        ("ACGTACGTACGT", [(M, 4), (I, 4), (M, 4)], 222),         # insertion whose anchor base is just left of the window
        ("ACGTACGTACGT", [(M, 5), (I, 4), (M, 3)], 222),         # insertion anchored on the first window column
        ("ACGTACGTACGT", [(M, 6), (I, 2), (M, 4)], 370),         # insertion anchored on the last window column
        ("ACGTNNGTACGT", [(EQ, 4), (X, 2), (N, 7), (M, 6)], 300),  # N bases, =/X ops, reference skip
        ("ACGTACGTACGT", [(M, 3), (P, 2), (M, 9)], 310),         # padding op: no case in the switch
        ("ACGTACGTACGT", [(M, 2), (I, 3), (D, 2), (M, 7)], 320), # insertion directly followed by a deletion
        ("ACGTACGTACGT", [(M, 12)], 100),                        # entirely left of the window
        ("ACGTACGTACGT", [(M, 12)], 376),                        # starts exactly at the window end
    ]
    rng = np.random.default_rng(7)
    quals = [[int(x) for x in rng.integers(0, 70, len(r[0]))] for r in reads]
    n = len(reads)
    sites.append(S([r[0] for r in reads], quals, [r[1] for r in reads], [r[2] for r in reads],
                   [int(x) for x in rng.integers(0, 100, n)], [1 if k % 2 else -1 for k in range(n)], [False] * n,
                   [k % 3 for k in range(n)], ref, 0, a0, a1, {"all": list(range(n)), "rev": list(range(n))[::-1], "none": []}))
    # long read over the whole window with many operations, both technologies mixed in one site
    long_read = "".join("ACGT"[(i * 5 + 1) % 4] for i in range(400))
    lq = [int(x) for x in rng.integers(0, 60, 400)]
    lc = [(SC, 10), (M, 60), (I, 12), (M, 40), (D, 7), (M, 55), (N, 3), (M, 70), (I, 1), (M, 90), (D, 1), (M, 52), (SC, 10)]
    sites.append(S([long_read, "ACGTACGTACGTACGT", long_read], [lq, [33] * 16, lq[::-1]],
                   [lc, [(M, 16)], lc], [120, 295, 131], [60, 3, 254], [1, -1, 1], [True, False, True], [1, 0, 2], ref, 0, 299, 310,
                   {"a": [0, 1, 2], "b": [1], "c": [2, 0]}))
    # window_start != 0 and an odd feature length is covered by the random sites below (window_start 1000)
    return sites


def main():
    from oracle import encoder_oracle as E, ref_encoder as R
    if not R.available():
        raise SystemExit("build oracle/_ref/libref_encoder.so first: make -C oracle")
    rng = np.random.default_rng(20241018)
    sites = handmade_sites(E)
    for k in range(24):
        sites.append(E.random_site(rng, n_reads=8, long_reads=(k % 3 == 2), border_cases=True, constant_quality=False,
                                   window_start=1000 if k % 2 else 0))
    queries, outputs = [], []
    for s, site in enumerate(sites):
        for allele in site.supports:
            for pac in (False, True):
                for hp in (False, True):
                    for L in ((150,) if s else (150, 151, 20)):
                        out = R.compute_features_colored_simple(site, allele, L, pac, hp)
                        queries.append((s, list(site.supports).index(allele), int(pac), int(hp), L))
                        outputs.append(out)
    obj = lambda xs: np.array(xs + [None], dtype=object)[:-1]
    flat = {}
    flat["n_sites"] = np.array(len(sites))
    for s, site in enumerate(sites):
        p = "s%d_" % s
        flat[p + "reads"] = np.array(site.reads)
        flat[p + "quals"] = np.concatenate([np.array(q, np.int32) for q in site.qualities])
        flat[p + "cig"] = np.concatenate([np.array(c, np.int32).reshape(-1, 2) for c in site.cigartuples])
        flat[p + "cig_n"] = np.array([len(c) for c in site.cigartuples], np.int32)
        flat[p + "per_read"] = np.array([site.reference_starts, site.mapq, site.orientation, [int(x) for x in site.pacbio], site.hp],
                                        np.int64)
        flat[p + "ref"] = np.array(site.reference)
        flat[p + "loc"] = np.array([site.window_start, site.assembly_start, site.assembly_stop], np.int64)
        flat[p + "alleles"] = np.array(list(site.supports))
        flat[p + "sup"] = np.concatenate([np.array(v, np.int64) for v in site.supports.values()] + [np.zeros(0, np.int64)])
        flat[p + "sup_n"] = np.array([len(v) for v in site.supports.values()], np.int32)
    flat["queries"] = np.array(queries, np.int64)
    flat["out_rows"] = np.array([o.shape[0] for o in outputs], np.int64)
    flat["out"] = np.concatenate([o.reshape(-1) for o in outputs])
    np.savez_compressed(OUT, **flat)
    # the restatement must agree before the file is worth committing
    bad = 0
    for (s, ai, pac, hp, L), want in zip(queries, outputs):
        got = E.compute_features_colored_simple(sites[s], list(sites[s].supports)[ai], L, bool(pac), bool(hp))
        bad += int(got.shape != want.shape or not np.array_equal(got, want))
    print("wrote", OUT, "sites", len(sites), "queries", len(queries), "rows", int(flat["out_rows"].sum()),
          "bytes", os.path.getsize(OUT), "oracle mismatches", bad)


if __name__ == "__main__":
    main()
