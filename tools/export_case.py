#!/usr/bin/env python
"""Write a case file for the plain-C host example (examples/c_host.c): packed weight blob + a synthetic ragged batch.

    python tools/export_case.py case.bin [--config single_tech] [--sites 64] [--coverage 20] [--precision bf16x3]
"""
import argparse
import ctypes as C
import os
import struct
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from hello_b200 import _lib, arch, synth, weights          # noqa: E402


def make_cfg(cfg, precision):
    c = _lib.HelloCfg()
    c.struct_size = C.sizeof(_lib.HelloCfg)
    c.n_tech = len(cfg.read_cin)
    for t, ch in enumerate(cfg.read_cin):
        c.read_channels[t] = ch
    for e in range(3):
        c.xattn_present[e] = int(cfg.xattn_present[e])
    c.has_combiners = _lib.COMBINE_SUM if cfg.legacy_sum else (_lib.COMBINE_CONV if cfg.combiners else _lib.COMBINE_NONE)
    c.meta_kind = {None: _lib.META_NONE, "meta_convolver": _lib.META_SITE, "meta_convolver_ref": _lib.META_REF}[cfg.meta]
    c.feature_length = arch.FEATURE_LENGTH
    c.precision = _lib.PRECISIONS[precision]
    c.max_chunk_sites = 0
    return c


def export(path, config="single_tech", sites=64, coverage=20, precision="bf16x3", seed=5, params=None):
    cfg = arch.CONFIGS[config]
    params = params or weights.init_params(cfg, seed=13)
    pl = synth.make_pileups(sites, coverage=coverage, channels=cfg.read_cin, seed=seed)
    blob = weights.pack_blob(cfg, params)
    R = [int(r.shape[0]) for r in pl.reads] + [0] * (2 - len(pl.reads))
    with open(path, "wb") as f:
        f.write(b"HELLOCAS")
        f.write(bytes(make_cfg(cfg, precision)))
        f.write(struct.pack("<5q", len(blob), pl.n_sites, pl.n_alleles, R[0], R[1]))
        f.write(blob)
        for r in pl.reads:
            f.write(r.contiguous().numpy().tobytes())
        for o in pl.allele_read_off:
            f.write(o.to("cpu").int().numpy().tobytes())
        f.write(pl.site_allele_off.int().numpy().tobytes())
        f.write(pl.ref_onehot.float().contiguous().numpy().tobytes())
    return cfg, params, pl


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("path")
    ap.add_argument("--config", default="single_tech", choices=sorted(arch.CONFIGS))
    ap.add_argument("--sites", type=int, default=64)
    ap.add_argument("--coverage", type=int, default=20)
    ap.add_argument("--precision", default="bf16x3", choices=sorted(_lib.PRECISIONS))
    a = ap.parse_args()
    export(a.path, a.config, a.sites, a.coverage, a.precision)
    print("wrote", a.path)
