"""ctypes binding of oracle/_ref/libref_encoder.so -- the reference's own C++ feature encoder
(/root/reference/c++/src/AlleleSearcherLiteFiltered.cpp:1031-1180), compiled by oracle/Makefile from the sources where
they lie.  TEST INFRASTRUCTURE: only tests/ and the golden-vector generators may import this file.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "_ref", "libref_encoder.so")
_lib = None


def available() -> bool:
    return os.path.exists(LIB_PATH)


def _load():
    global _lib
    if _lib is None:
        lib = C.CDLL(LIB_PATH)
        lib.ref_encode_allele.restype = C.c_long
        lib.ref_encode_allele.argtypes = [
            C.c_int, C.c_char_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
            C.c_void_p, C.c_void_p, C.c_void_p, C.c_char_p, C.c_long, C.c_long, C.c_long, C.c_void_p, C.c_int, C.c_int,
            C.c_int, C.c_int, C.c_void_p, C.c_long]
        _lib = lib
    return _lib


def compute_features_colored_simple(site, allele: str, feature_length: int, pacbio_: bool, include_hp_tags: bool) -> np.ndarray:
    """Same signature as oracle.encoder_oracle.compute_features_colored_simple, answered by the compiled reference."""
    lib = _load()
    n = len(site.reads)
    bases = "".join(site.reads).encode()
    quals = np.array([q for qs in site.qualities for q in qs], np.uint8)
    read_off = np.zeros(n + 1, np.int64)
    np.cumsum([len(r) for r in site.reads], out=read_off[1:])
    assert quals.size == read_off[-1]
    ops = np.array([op for c in site.cigartuples for op, _ in c], np.int32)
    lens = np.array([ln for c in site.cigartuples for _, ln in c], np.int32)
    cig_off = np.zeros(n + 1, np.int64)
    np.cumsum([len(c) for c in site.cigartuples], out=cig_off[1:])
    starts = np.array(site.reference_starts, np.int64)
    mapq = np.array(site.mapq, np.int32)
    orient = np.array(site.orientation, np.int32)
    pacbio = np.array([int(bool(x)) for x in site.pacbio], np.int32)
    hp = np.array(site.hp, np.int32)
    ids = np.array(site.supports.get(allele, []), np.int64)
    n_ch = 7 if include_hp_tags else 6
    cap = max(1, ids.size)
    out = np.zeros((cap, feature_length, n_ch), np.uint8)
    ptr = lambda a: a.ctypes.data_as(C.c_void_p)
    rows = lib.ref_encode_allele(n, bases, ptr(quals), ptr(read_off), ptr(ops), ptr(lens), ptr(cig_off), ptr(starts),
                                 ptr(mapq), ptr(orient), ptr(pacbio), ptr(hp), site.reference.encode(), site.window_start,
                                 site.assembly_start, site.assembly_stop, ptr(ids), int(ids.size), feature_length,
                                 int(bool(pacbio_)), int(bool(include_hp_tags)), ptr(out), cap)
    if rows < 0:
        raise RuntimeError("ref_encode_allele failed (%d)" % rows)
    return out[:rows].copy()
