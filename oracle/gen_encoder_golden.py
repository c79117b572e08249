"""Generate tests/golden/encoder.npz from the reference's own Python specification of the read encoding.

Run in the build container only:   PYTHONDONTWRITEBYTECODE=1 python oracle/gen_encoder_golden.py

``/root/reference/python/test_aligner.py:create_read_encoding`` (:108-183) is what the reference's test asserts the
C++ ``computeFeaturesColoredSimple`` against.  It is imported unmodified (its ``import libCallability`` -- the
Boost.Python module that cannot be built here -- is satisfied by an empty stub; the specification never calls it) and
run on
  * the two fixture pileups of the reference's test (tagless and haplotagged, test_aligner.py:300-380), and
  * seeded random reads with matches / insertions / deletions / skips / clips, restricted to the domain where the
    specification and the C++ agree (see oracle/encoder_oracle.py: constant quality per read, no deletion straddling a
    window border or opening a read, no soft clips -- the specification has no branch for them --, bases in ACGT).
Stored per read: the inputs and the specification's [L, C] encoding.
"""
from __future__ import annotations

import os
import sys
import types

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference/python"
OUT = os.path.join(ROOT, "tests", "golden", "encoder.npz")
sys.path.insert(0, ROOT)


def in_agreement_domain(cigar, ref_start, start, end):
    from oracle.encoder_oracle import BAM_CDEL, BAM_CINS, BAM_CSOFT_CLIP, BAM_CMATCH, BAM_CEQUAL, BAM_CDIFF, BAM_CREF_SKIP
    rf, rd = ref_start, 0
    for op, n in cigar:
        if op in (BAM_CMATCH, BAM_CEQUAL, BAM_CDIFF):
            rf += n; rd += n
        elif op == BAM_CDEL:
            if rd == 0:
                return False
            if not (rf < start or rf - 1 >= end or (rf - 1 >= start and rf + n - 1 < end)):
                return False
            rf += n
        elif op == BAM_CREF_SKIP:
            rf += n
        elif op == BAM_CSOFT_CLIP:
            return False                 # the specification has no soft-clip branch (its read pointer does not move)
        elif op == BAM_CINS:
            rd += n
    return True


def main():
    sys.modules["libCallability"] = types.ModuleType("libCallability")
    sys.path.insert(0, REF)
    import test_aligner as T                                   # the reference's specification
    from oracle import encoder_oracle as E

    cases = []   # (reference, window_start(=0), variant_range, feature_length, read, qual, cigar, ref_start, mapq, orient, hp or -1)
    # the reference's fixtures (test_aligner.py:300-380); differing region of the sketch in its comment: [10, 14)
    reference = "ACGATACCGTACGGATCGGATCGT"
    fx = [("TAATCG", [26] * 6, [[0, 2], [2, 3], [0, 4]], 9, 30, -1), ("TAACGGATCG", [30] * 10, [[0, 2], [1, 1], [0, 7]], 9, 44, 1),
          ("TGCGGATCG", [15] * 9, [[0, 9]], 9, 75, 1)]
    for hps in (None, (1, 0, 2)):
        for k, (read, q, cg, rs, mq, ori) in enumerate(fx):
            cases.append((reference, (10, 14), 10, read, q, cg, rs, mq, ori, -1 if hps is None else hps[k]))
    rng = np.random.default_rng(20240611)
    n_random = 0
    while n_random < 60:
        long_reads = n_random % 3 == 2
        site = E.random_site(rng, n_reads=6, long_reads=long_reads, border_cases=True, constant_quality=True, window_start=0)
        L = 150
        mid = (site.assembly_start + site.assembly_stop) // 2
        start, end = mid - L // 2, mid - L // 2 + L
        for r in range(len(site.reads)):
            if "N" in site.reads[r] or not in_agreement_domain(site.cigartuples[r], site.reference_starts[r], start, end):
                continue
            hp = site.hp[r] if n_random % 2 else -1
            cases.append((site.reference, (site.assembly_start, site.assembly_stop), L, site.reads[r], site.qualities[r],
                          [list(c) for c in site.cigartuples[r]], site.reference_starts[r], site.mapq[r], site.orientation[r], hp))
            n_random += 1
    enc, meta = [], []
    for (reference, vr, L, read, q, cg, rs, mq, ori, hp) in cases:
        desc = T.ReadDescriptor(read=read, name=0, quality=q, cigartuples=cg, reference_start=rs, mapq=mq, orientation=ori,
                                pacbio=False, hp=None if hp < 0 else hp)
        arr, _ = T.create_read_encoding(desc, reference, feature_length=L, variant_range=vr)
        assert arr.min() >= 0 and arr.max() <= 255
        enc.append(arr.T.astype(np.uint8))
        meta.append((vr[0], vr[1], L, rs, mq, ori, hp))
    obj = lambda xs: np.array(xs, dtype=object)
    np.savez_compressed(OUT, references=obj([c[0] for c in cases]), reads=obj([c[3] for c in cases]),
                        quals=obj([np.array(c[4], np.int32) for c in cases]),
                        cigars=obj([np.array(c[5], np.int32).reshape(-1, 2) for c in cases]),
                        meta=np.array(meta, np.int64), encodings=obj(enc))
    print("wrote", OUT, "reads", len(cases), "(6 fixture + %d random)" % (len(cases) - 6))


if __name__ == "__main__":
    main()
