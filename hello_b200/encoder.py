"""Host side of the GPU read-feature encoder (include/hello_encode.h).

Mirrors the reference interface for this step: an ``AlleleSearcherLite`` object is built per site from the aligned
reads (``libCallability.AlleleSearcherLite(reads, names, qualities, cigartuples, referenceStarts, mapq, orientation,
pacbio, hp, reference, windowStart, ...)``, python/test_aligner.py:222-238) and asked, per allele and technology, for
``computeFeaturesColoredSimple(allele, featureLength, pacbio_, include_hp_tags)`` (c++/src/
AlleleSearcherLiteFiltered.cpp:1031-1180).  ``SiteEncoderB200`` answers that call for one site; ``encode_sites`` encodes a
whole batch of sites in ONE launch straight into the CSR tensors ``MoEEngine.run`` takes, which is the form the
product path uses (the pileups never exist as 900-byte rows in host memory).

There is no CPU fallback: everything here ends in ``hello_encode_reads`` of libhello_moe.so.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch

from . import _lib

FEATURE_LENGTH = 150


@dataclass
class SitePileup:
    """The per-site state of the reference's AlleleSearcherLiteFiltered that the encoder reads."""
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
    supports: Dict[str, List[int]] = field(default_factory=dict)


class HelloEncodeBatch(C.Structure):
    _fields_ = [
        ("n_rows", C.c_int64), ("feature_length", C.c_int32), ("channels", C.c_int32),
        ("d_row_read", C.c_void_p), ("d_row_site", C.c_void_p),
        ("d_read_off", C.c_void_p), ("d_bases", C.c_void_p), ("d_quals", C.c_void_p),
        ("d_cigar_off", C.c_void_p), ("d_cigars", C.c_void_p), ("d_ref_start", C.c_void_p),
        ("d_mapq", C.c_void_p), ("d_orientation", C.c_void_p), ("d_hp", C.c_void_p),
        ("d_ref_off", C.c_void_p), ("d_reference", C.c_void_p), ("d_window_start", C.c_void_p),
        ("d_assembly_start", C.c_void_p), ("d_assembly_stop", C.c_void_p),
    ]


def _load():
    lib = _lib.load()
    if not getattr(lib, "_encode_bound", False):
        lib.hello_encode_reads.restype = C.c_int
        lib.hello_encode_reads.argtypes = [C.POINTER(HelloEncodeBatch), C.c_void_p, C.c_void_p]
        lib.hello_encode_last_error.restype = C.c_char_p
        lib.hello_encode_last_error.argtypes = []
        lib._encode_bound = True
    return lib


@dataclass
class PackedReads:
    """Aligned reads and reference windows of a batch of sites as flat host arrays (what crosses PCIe)."""
    read_off: np.ndarray       # int64 [n_reads+1]
    bases: np.ndarray          # uint8
    quals: np.ndarray          # uint8
    cigar_off: np.ndarray      # int64 [n_reads+1]
    cigars: np.ndarray         # uint32, BAM encoding
    ref_start: np.ndarray      # int64 [n_reads]
    mapq: np.ndarray           # uint8
    orientation: np.ndarray    # int8
    hp: np.ndarray             # uint8
    read_base: np.ndarray      # int64 [n_sites+1]: first global read index of every site
    ref_off: np.ndarray        # int64 [n_sites+1]
    reference: np.ndarray      # uint8
    window_start: np.ndarray   # int64 [n_sites]
    assembly_start: np.ndarray
    assembly_stop: np.ndarray

    def nbytes(self) -> int:
        return sum(getattr(self, f).nbytes for f in self.__dataclass_fields__)


def pack_sites(sites: Sequence) -> PackedReads:
    n_reads = sum(len(s.reads) for s in sites)
    read_off = np.zeros(n_reads + 1, np.int64)
    cigar_off = np.zeros(n_reads + 1, np.int64)
    bases, quals, cigars = [], [], []
    ref_start = np.zeros(n_reads, np.int64)
    mapq = np.zeros(n_reads, np.uint8)
    orient = np.zeros(n_reads, np.int8)
    hp = np.zeros(n_reads, np.uint8)
    read_base = np.zeros(len(sites) + 1, np.int64)
    ref_off = np.zeros(len(sites) + 1, np.int64)
    refs = []
    r = 0
    for si, s in enumerate(sites):
        read_base[si] = r
        for k in range(len(s.reads)):
            b = np.frombuffer(s.reads[k].encode("ascii"), np.uint8)
            q = np.asarray(s.qualities[k], np.int64)
            if q.size != b.size:
                raise ValueError("site %d read %d: %d qualities for %d bases" % (si, k, q.size, b.size))
            cg = np.asarray(s.cigartuples[k], np.int64).reshape(-1, 2)
            if (cg[:, 0] < 0).any() or (cg[:, 0] > 9).any() or (cg[:, 1] < 0).any() or (cg[:, 1] >= 1 << 28).any():
                raise ValueError("site %d read %d: bad CIGAR" % (si, k))
            consumed = int(cg[np.isin(cg[:, 0], (0, 1, 4, 7, 8)), 1].sum())
            if consumed > b.size:
                raise ValueError("site %d read %d: CIGAR consumes %d bases, read has %d" % (si, k, consumed, b.size))
            bases.append(b); quals.append(np.clip(q, 0, 255).astype(np.uint8))
            cigars.append((cg[:, 1] << 4 | cg[:, 0]).astype(np.uint32))
            read_off[r + 1] = read_off[r] + b.size
            cigar_off[r + 1] = cigar_off[r] + cg.shape[0]
            ref_start[r] = s.reference_starts[k]
            mapq[r] = min(max(int(s.mapq[k]), 0), 255)
            orient[r] = 1 if s.orientation[k] > 0 else -1
            hp[r] = min(max(int(s.hp[k]), 0), 255)
            r += 1
        refs.append(np.frombuffer(s.reference.encode("ascii"), np.uint8))
        ref_off[si + 1] = ref_off[si] + len(s.reference)
    read_base[len(sites)] = r
    cat = lambda xs, dt: np.concatenate(xs).astype(dt) if xs else np.zeros(0, dt)
    return PackedReads(read_off, cat(bases, np.uint8), cat(quals, np.uint8), cigar_off, cat(cigars, np.uint32), ref_start,
                       mapq, orient, hp, read_base, ref_off, cat(refs, np.uint8),
                       np.array([s.window_start for s in sites], np.int64),
                       np.array([s.assembly_start for s in sites], np.int64),
                       np.array([s.assembly_stop for s in sites], np.int64))


def row_plan(sites: Sequence, alleles_per_site: Sequence[Sequence[str]], pacbio: bool, read_base: np.ndarray):
    """Rows in network order for one technology: site -> allele (given order) -> supporting reads of that technology
    (numReadsSupportingAlleleStrict, :958-968), one -1 row for an allele without support (:1037-1043).
    Returns (row_read int32, row_site int32, reads-per-allele counts)."""
    row_read, row_site, counts = [], [], []
    for si, (s, alleles) in enumerate(zip(sites, alleles_per_site)):
        for a in alleles:
            ids = [r for r in s.supports.get(a, []) if bool(s.pacbio[r]) == bool(pacbio)]
            if ids:
                row_read.extend(int(read_base[si]) + r for r in ids)
                row_site.extend([si] * len(ids))
                counts.append(len(ids))
            else:
                row_read.append(-1); row_site.append(si); counts.append(1)
    return np.asarray(row_read, np.int32), np.asarray(row_site, np.int32), counts


class DevicePackedReads:
    """PackedReads uploaded once; several row plans (technologies) can be encoded from it."""

    def __init__(self, packed: PackedReads, device="cuda:0", non_blocking: bool = True):
        self.device = torch.device(device)
        if self.device.type != "cuda" or not torch.cuda.is_available():
            raise _lib.HelloMoEError("hello_b200 needs a CUDA device (sm_100a); there is no CPU fallback")
        self.packed = packed
        up = lambda a: torch.from_numpy(a.view(np.int32) if a.dtype == np.uint32 else a).to(self.device, non_blocking=non_blocking)
        self.t = {f: up(getattr(packed, f)) for f in packed.__dataclass_fields__ if f != "read_base"}

    def encode(self, row_read, row_site, channels: int = 6,
               feature_length: int = FEATURE_LENGTH, out: Optional[torch.Tensor] = None) -> torch.Tensor:
        """uint8 [n_rows, feature_length, channels] on the device (asynchronous, on the current stream)."""
        lib = _load()
        as_dev = lambda a: a if torch.is_tensor(a) else torch.from_numpy(np.ascontiguousarray(a, np.int32))
        rr = as_dev(row_read).to(self.device, non_blocking=True)      # device tensors are used as they are
        rs = as_dev(row_site).to(self.device, non_blocking=True)
        if rr.dtype != torch.int32 or rs.dtype != torch.int32 or rr.numel() != rs.numel():
            raise ValueError("row_read / row_site must be int32 arrays of equal length")
        n = int(rr.numel())
        if out is None:
            out = torch.empty((n, feature_length, channels), dtype=torch.uint8, device=self.device)
        b = HelloEncodeBatch()
        b.n_rows, b.feature_length, b.channels = n, feature_length, channels
        b.d_row_read, b.d_row_site = rr.data_ptr(), rs.data_ptr()
        t = self.t
        b.d_read_off, b.d_bases, b.d_quals = t["read_off"].data_ptr(), t["bases"].data_ptr(), t["quals"].data_ptr()
        b.d_cigar_off, b.d_cigars, b.d_ref_start = t["cigar_off"].data_ptr(), t["cigars"].data_ptr(), t["ref_start"].data_ptr()
        b.d_mapq, b.d_orientation, b.d_hp = t["mapq"].data_ptr(), t["orientation"].data_ptr(), t["hp"].data_ptr()
        b.d_ref_off, b.d_reference = t["ref_off"].data_ptr(), t["reference"].data_ptr()
        b.d_window_start = t["window_start"].data_ptr()
        b.d_assembly_start, b.d_assembly_stop = t["assembly_start"].data_ptr(), t["assembly_stop"].data_ptr()
        stream = torch.cuda.current_stream(self.device).cuda_stream
        with torch.cuda.device(self.device):
            rc = lib.hello_encode_reads(C.byref(b), out.data_ptr(), C.c_void_p(stream))
        if rc != 0:
            raise _lib.HelloMoEError("hello_encode_reads failed (%d): %s" % (rc, lib.hello_encode_last_error().decode()))
        self._keep = (rr, rs)
        return out


def encode_sites(sites: Sequence, alleles_per_site: Sequence[Sequence[str]], technologies: Sequence[Tuple[bool, int]] = ((False, 6),),
                 device="cuda:0", feature_length: int = FEATURE_LENGTH):
    """Encode a batch of sites for the network: returns (reads per technology [R_t, L, C_t] uint8 on the device,
    allele_read_off per technology int32 [A+1], site_allele_off int32 [S+1]) -- the arguments of DeviceBatch.
    `technologies`: (pacbio flag, channels) per network input, e.g. ((False, 6), (True, 6)) for a hybrid model."""
    packed = pack_sites(sites)
    dev = DevicePackedReads(packed, device)
    reads, offs = [], []
    for pacbio, ch in technologies:
        rr, rs, counts = row_plan(sites, alleles_per_site, pacbio, packed.read_base)
        reads.append(dev.encode(rr, rs, ch, feature_length))
        off = np.zeros(len(counts) + 1, np.int64)
        np.cumsum(counts, out=off[1:])
        offs.append(torch.from_numpy(off.astype(np.int32)))
    sao = np.zeros(len(sites) + 1, np.int64)
    np.cumsum([len(a) for a in alleles_per_site], out=sao[1:])
    return tuple(reads), tuple(offs), torch.from_numpy(sao.astype(np.int32))


class SiteEncoderB200:
    """One site, asked like the reference's searcher object: ``computeFeaturesColoredSimple(allele, featureLength,
    pacbio_, include_hp_tags)`` -> uint8 numpy array [n, featureLength, channels]."""

    def __init__(self, site, device="cuda:0"):
        self.site = site
        self.packed = pack_sites([site])
        self.dev = DevicePackedReads(self.packed, device)

    def computeFeaturesColoredSimple(self, allele: str, featureLength: int, pacbio_: bool, include_hp_tags: bool) -> np.ndarray:
        rr, rs, _ = row_plan([self.site], [[allele]], pacbio_, self.packed.read_base)
        out = self.dev.encode(rr, rs, 7 if include_hp_tags else 6, featureLength)
        return out.cpu().numpy()
