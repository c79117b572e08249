"""CPU oracle for the read-feature encoder -- TEST INFRASTRUCTURE, NOT A PRODUCT PATH.

Restates ``AlleleSearcherLiteFiltered::computeFeaturesColoredSimple`` (/root/reference/c++/src/
AlleleSearcherLiteFiltered.cpp:1031-1180) and its colour functions (:971-1027, constants :360-384) in plain Python
loops over numpy arrays: one uint8 row [L, C] per supporting read, walked CIGAR operation by operation exactly as the
C++ does (including its switch fall-throughs: a deletion is only drawn when the base BEFORE it lies in the window,
an insertion colours the base before it with the minimum quality of that base and the inserted bases).
Only tests/ and tools/bench_encoder.py's cpu_baseline leg may import this file.

Parity pin (two independent ones):
  * the reference's COMPILED C++: oracle/Makefile builds /root/reference/c++/src/*.cpp where they lie (against
    oracle/boost_shim, a stand-in for the Boost.Python container types -- Boost is not in this image and the reference's
    CMake build, with its hard-coded /opt/boost and python3.7m, is not run) into oracle/_ref/libref_encoder.so;
    tests/golden/encoder_cpp.npz (oracle/gen_encoder_cpp_golden.py) holds its outputs for 26 sites / 432 queries covering
    clips, indels at the window borders and at the start of a read, insertions into reads of varying quality, N bases,
    the technology filter and the no-support row, and tests/test_encoder_oracle.py compares fresh random sites live when
    the library is present.  This restatement is bit-identical on all of them.
  * the reference's own Python specification of the encoding, ``python/test_aligner.py:create_read_encoding`` (:108-183)
    -- the function the reference's test asserts the C++ against -- on the reference's two fixture pileups (:300-380)
    and on seeded random reads (tests/golden/encoder.npz, oracle/gen_encoder_golden.py).  The specification and the C++
    differ at the window borders, for insertions into reads of varying quality and for soft clips (the specification has
    no branch for them); those golden cases stay inside the domain where the two agree.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np

BAM_CMATCH, BAM_CINS, BAM_CDEL, BAM_CREF_SKIP, BAM_CSOFT_CLIP, BAM_CHARD_CLIP, BAM_CPAD, BAM_CEQUAL, BAM_CDIFF = range(9)
READ_BASE, REF_BASE, READ_QUAL, READ_MAPQ, READ_ORIENT, POSITION_MARKER, HP = range(7)


def base_color(base: str) -> int:                       # :971-984
    return {"A": 40 + 3 * 70, "G": 40 + 2 * 70, "T": 30 + 1 * 70, "C": 30}.get(base, 0)


def base_quality_color(qual: int) -> int:               # :987-991  (float capped; double arithmetic; truncation)
    capped = float(np.float32(min(qual, 40)))
    return int(254 * (1.0 * capped / 40))


def mapping_quality_color(qual: int) -> int:            # :994-998
    capped = float(np.float32(min(qual, 60)))
    return int(254 * (1.0 * capped / 60))


def strand_color(value: int) -> int:                    # :1001-1004
    return 70 if value > 0 else 240


def hp_color(hp: int) -> int:                           # :1018-1027
    return 120 if hp == 1 else (240 if hp == 2 else 0)


@dataclass
class SitePileup:
    """What one AlleleSearcherLiteFiltered holds for a site (constructor :330-360)."""
    reads: List[str]
    qualities: List[Sequence[int]]
    cigartuples: List[Sequence[Tuple[int, int]]]
    reference_starts: List[int]
    mapq: List[int]
    orientation: List[int]
    pacbio: List[bool]
    hp: List[int]
    reference: str
    window_start: int
    assembly_start: int
    assembly_stop: int
    supports: Dict[str, List[int]] = field(default_factory=dict)     # allele -> read ids, in iteration order


def compute_features_colored_simple(site: SitePileup, allele: str, feature_length: int, pacbio_: bool,
                                    include_hp_tags: bool) -> np.ndarray:
    """uint8 [n, feature_length, channels]; one all-zero row when the allele has no support in this technology."""
    n_ch = 7 if include_hp_tags else 6
    ids = [r for r in site.supports.get(allele, []) if bool(site.pacbio[r]) == bool(pacbio_)]
    if not ids:
        return np.zeros((1, feature_length, n_ch), np.uint8)
    out = np.zeros((len(ids), feature_length, n_ch), np.uint8)
    mid = (site.assembly_start + site.assembly_stop) // 2
    start = mid - feature_length // 2
    end = start + feature_length

    def position_color(position: int) -> int:           # :1007-1015
        return 240 if site.assembly_start - site.window_start <= position < site.assembly_stop - site.window_start else 70

    for row, rid in enumerate(ids):
        read, qual = site.reads[rid], site.qualities[rid]
        rf, rd = site.reference_starts[rid], 0
        mq, sc, hc = mapping_quality_color(site.mapq[rid]), strand_color(site.orientation[rid]), hp_color(site.hp[rid])
        a = out[row]
        for op, length in site.cigartuples[rid]:
            if op in (BAM_CEQUAL, BAM_CDIFF, BAM_CMATCH):
                for j in range(length):
                    if start <= rf + j < end:
                        f = rf + j - start
                        a[f, READ_BASE] = base_color(read[rd + j])
                        a[f, REF_BASE] = base_color(site.reference[rf + j - site.window_start])
                        a[f, READ_QUAL] = base_quality_color(qual[rd + j])
                        a[f, READ_MAPQ], a[f, READ_ORIENT] = mq, sc
                        a[f, POSITION_MARKER] = position_color(rf + j - site.window_start)
                        if include_hp_tags:
                            a[f, HP] = hc
                rf += length
                rd += length
            elif op in (BAM_CDEL, BAM_CREF_SKIP):
                if op == BAM_CDEL and start <= rf - 1 < end:
                    for i in range(rf - 1, rf + length):
                        if not (start <= i < end):
                            continue
                        f = i - start
                        a[f, REF_BASE] = base_color(site.reference[i - site.window_start])
                        a[f, READ_MAPQ], a[f, READ_ORIENT] = mq, sc
                        a[f, POSITION_MARKER] = position_color(i - site.window_start)
                        if include_hp_tags:
                            a[f, HP] = hc
                    f = rf - 1 - start
                    a[f, READ_BASE] = base_color("*")
                    a[f, READ_QUAL] = base_quality_color(qual[rd - 1]) if rd > 0 else 0
                rf += length                            # BAM_CDEL falls through into BAM_CREF_SKIP
            elif op in (BAM_CINS, BAM_CSOFT_CLIP):
                if op == BAM_CINS and start <= rf - 1 < end:
                    f = rf - 1 - start
                    lo = rd - 1 if rd > 0 else rd
                    a[f, READ_BASE] = base_color("*")
                    a[f, REF_BASE] = base_color(site.reference[rf - 1 - site.window_start])
                    a[f, READ_QUAL] = base_quality_color(min(qual[lo:rd + length]))
                    a[f, READ_MAPQ], a[f, READ_ORIENT] = mq, sc
                    a[f, POSITION_MARKER] = position_color(rf - 1 - site.window_start)
                    if include_hp_tags:
                        a[f, HP] = hc
                rd += length                            # BAM_CINS falls through into BAM_CSOFT_CLIP
            # BAM_CHARD_CLIP, BAM_CPAD, BAM_CBACK: no case in the switch
    return out


def random_site(rng: np.random.Generator, n_reads: int = 12, feature_length: int = 150, long_reads: bool = False,
                border_cases: bool = True, constant_quality: bool = False, window_start: int = 1000) -> SitePileup:
    """A synthetic site with reads carrying matches, mismatches, insertions, deletions, skips and clips.  With
    `border_cases` the reads may start / end / carry indels anywhere relative to the feature window."""
    ref_len = 700 if long_reads else 450
    reference = "".join(rng.choice(list("ACGT"), ref_len))
    a0 = window_start + ref_len // 2 + int(rng.integers(-20, 20))
    a1 = a0 + int(rng.integers(1, 12))
    mid = (a0 + a1) // 2
    start = mid - feature_length // 2
    site = SitePileup([], [], [], [], [], [], [], [], reference, window_start, a0, a1, {})
    for r in range(n_reads):
        read_len = int(rng.integers(250, 420)) if long_reads else int(rng.integers(90, 160))
        if border_cases:
            rs = start + int(rng.integers(-read_len + 5, feature_length - 5))
        else:
            rs = start + int(rng.integers(-20, 5))
        rs = max(window_start + 2, rs)
        ops, rf, used = [], rs, 0
        if rng.random() < 0.3:
            k = int(rng.integers(1, 8)); ops.append((BAM_CSOFT_CLIP, k)); used += k
        if rng.random() < 0.1:
            ops.append((BAM_CHARD_CLIP, int(rng.integers(1, 5))))
        while used < read_len - 12 and rf < window_start + ref_len - 40:
            k = int(rng.integers(3, 40))
            k = min(k, read_len - used - 6, window_start + ref_len - 30 - rf)
            if k <= 0:
                break
            ops.append((int(rng.choice([BAM_CMATCH, BAM_CEQUAL, BAM_CDIFF])), k)); rf += k; used += k
            u = rng.random()
            if u < 0.25:
                k = int(rng.integers(1, 30 if long_reads else 6)); k = min(k, read_len - used - 4)
                if k > 0:
                    ops.append((BAM_CINS, k)); used += k
            elif u < 0.5:
                k = int(rng.integers(1, 9)); ops.append((BAM_CDEL, k)); rf += k
            elif u < 0.55:
                k = int(rng.integers(1, 6)); ops.append((BAM_CREF_SKIP, k)); rf += k
        if rng.random() < 0.3 and read_len - used > 2:
            ops.append((BAM_CSOFT_CLIP, read_len - used)); used = read_len
        seq = "".join(rng.choice(list("ACGTN"), used, p=[.245, .245, .245, .245, .02]))
        q = [int(rng.integers(2, 60))] * used if constant_quality else [int(x) for x in rng.integers(0, 60, used)]
        site.reads.append(seq); site.qualities.append(q); site.cigartuples.append(ops)
        site.reference_starts.append(rs); site.mapq.append(int(rng.integers(0, 90)))
        site.orientation.append(int(rng.choice([-1, 1]))); site.pacbio.append(bool(long_reads))
        site.hp.append(int(rng.integers(0, 3)))
    ids = list(rng.permutation(n_reads))
    cut = sorted(rng.choice(np.arange(1, n_reads), size=min(2, n_reads - 1), replace=False)) if n_reads > 1 else []
    names = ["ref", "alt1", "alt2"]
    bounds = [0] + [int(c) for c in cut] + [n_reads]
    for k in range(len(bounds) - 1):
        site.supports[names[k]] = [int(x) for x in ids[bounds[k]:bounds[k + 1]]]
    site.supports["unsupported"] = []
    return site
