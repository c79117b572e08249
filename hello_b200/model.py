"""Host side of the B200 MoE forward: the reference's two call signatures on top of the C ABI.

  * ``MoEAttentionB200.forward(tensors, numAllelesPerSite, numReadsPerAllele, reference_segments, *args)``
    -- same arguments and return convention as ``MoEAttention.forward``
    (reference: python/MixtureOfExpertsAdvanced.py:161-252), the batched call ``WrapperForDataParallel`` makes
    (python/MixtureOfExpertsDNNFast.py:128-134);
  * ``MoEMergedWrapperB200(featureDict, segment)`` -- same as ``MoEMergedWrapperAdvanced.forward`` (:520-589),
    the per-site call of python/caller_calling.py:651-652, with ``.eval()`` and ``.providePredictions``.

PyTorch is used for device memory, streams and host<->device copies only; all arithmetic runs in
libhello_moe.so (hand-written sm_100a CUDA).  There is no CPU path: a missing library or a missing GPU raises.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch

from . import _lib, arch, weights
from .synth import Pileups, pair_offsets


def _csr(counts) -> torch.Tensor:
    counts = np.asarray(counts, dtype=np.int64)
    if counts.ndim != 1 or (counts < 1).any():
        raise ValueError("every CSR slot must hold at least one row (reduceSlots requires it)")
    off = np.zeros(counts.size + 1, dtype=np.int64)
    np.cumsum(counts, out=off[1:])
    if off[-1] >= 2 ** 31:
        raise ValueError("CSR offsets exceed int32")
    return torch.from_numpy(off.astype(np.int32))


@dataclass
class DeviceBatch:
    """A ragged batch resident on the GPU, in the layout hello_moe_forward takes."""
    reads: Tuple[torch.Tensor, ...]            # uint8, device
    layout: int
    allele_read_off_h: Tuple[torch.Tensor, ...]
    allele_read_off_d: Tuple[torch.Tensor, ...]
    site_allele_off_h: torch.Tensor
    site_allele_off_d: torch.Tensor
    pair_off_h: torch.Tensor
    pair_off_d: torch.Tensor
    ref_onehot: Optional[torch.Tensor] = None  # fp32 [S, L, 5], device
    allele_rank: Optional[torch.Tensor] = None  # int32 [A], device

    @property
    def n_sites(self) -> int:
        return self.site_allele_off_h.numel() - 1

    @property
    def n_alleles(self) -> int:
        return int(self.site_allele_off_h[-1])

    @property
    def n_pairs(self) -> int:
        return int(self.pair_off_h[-1])

    def input_bytes(self) -> int:
        n = sum(r.numel() for r in self.reads)
        n += sum(o.numel() * 4 for o in self.allele_read_off_h) + self.site_allele_off_h.numel() * 4
        n += self.pair_off_h.numel() * 8
        if self.ref_onehot is not None:
            n += self.ref_onehot.numel() * 4
        return n

    @staticmethod
    def from_host(reads: Sequence[torch.Tensor], layout: int, allele_read_off: Sequence[torch.Tensor],
                  site_allele_off: torch.Tensor, ref_onehot: Optional[torch.Tensor], device,
                  allele_rank: Optional[torch.Tensor] = None, non_blocking: bool = True,
                  pin: bool = False) -> "DeviceBatch":
        dev = torch.device(device)
        pair_off = pair_offsets(site_allele_off)
        if pin:
            pair_off = pair_off.pin_memory()
        up = lambda t: t.to(dev, non_blocking=non_blocking)
        return DeviceBatch(
            reads=tuple(up(r.contiguous()) for r in reads), layout=layout,
            allele_read_off_h=tuple(o.contiguous() for o in allele_read_off),
            allele_read_off_d=tuple(up(o.contiguous()) for o in allele_read_off),
            site_allele_off_h=site_allele_off.contiguous(), site_allele_off_d=up(site_allele_off.contiguous()),
            pair_off_h=pair_off, pair_off_d=up(pair_off),
            ref_onehot=up(ref_onehot.contiguous().float()) if ref_onehot is not None else None,
            allele_rank=up(allele_rank.contiguous()) if allele_rank is not None else None)

    @staticmethod
    def from_pileups(pl: Pileups, device, need_ref: bool = True) -> "DeviceBatch":
        return DeviceBatch.from_host(pl.reads, _lib.LAYOUT_RLC, pl.allele_read_off, pl.site_allele_off,
                                     pl.ref_onehot if need_ref else None, device)


@dataclass
class BatchResult:
    logits: torch.Tensor       # [3, A]
    meta: torch.Tensor         # [S, 3]
    pair_prob: torch.Tensor    # [4, P]   mixed, P_e0, P_e1, P_e2
    pair_mix64: torch.Tensor   # [P]      float64 re-mix (prepareVcf.py:154-162)
    best_pair: torch.Tensor    # [S, 2]
    best_prob: torch.Tensor    # [S]
    pair_off: torch.Tensor     # [S+1] host
    # the final-call step (prepareVcf.py:36-105,142-175): call 0 = fp32 mixture, 1-3 = experts, 4 = float64 re-mix
    call_pair: torch.Tensor    # [S, 5, 2] int32
    call_qual: torch.Tensor    # [S, 5]    float64  QUAL
    best_expert: torch.Tensor  # [S]       int32    np.argmax(meta)

    def tensors(self):
        return (self.logits, self.meta, self.pair_prob, self.pair_mix64, self.best_pair, self.best_prob,
                self.call_pair, self.call_qual, self.best_expert)

    def output_bytes(self) -> int:
        return sum(t.numel() * t.element_size() for t in self.tensors())


class MoEEngine:
    """Owns one hello_moe handle (weights on one GPU) plus its workspace."""

    def __init__(self, cfg: arch.ModelConfig, params: Dict[str, torch.Tensor], device="cuda:0",
                 precision: str = "bf16x3", workspace_bytes: int = 4 << 30, max_chunk_sites: int = 0):
        if not torch.cuda.is_available():
            raise _lib.HelloMoEError("hello_b200 needs a CUDA device (sm_100a); there is no CPU fallback")
        self.lib = _lib.load()
        self.cfg = cfg
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise _lib.HelloMoEError("hello_b200 runs on CUDA devices only")
        self.precision = precision
        index = self.device.index if self.device.index is not None else torch.cuda.current_device()
        self.device = torch.device("cuda", index)
        blob = weights.pack_blob(cfg, params)
        c = _lib.HelloCfg()
        c.struct_size = C.sizeof(_lib.HelloCfg)
        c.n_tech = len(cfg.read_cin)
        for t, ch in enumerate(cfg.read_cin):
            c.read_channels[t] = ch
        for e in range(3):
            c.xattn_present[e] = int(cfg.xattn_present[e])
        c.has_combiners = _lib.COMBINE_SUM if cfg.legacy_sum else (_lib.COMBINE_CONV if cfg.combiners else _lib.COMBINE_NONE)
        c.meta_kind = {None: _lib.META_NONE, "meta_convolver": _lib.META_SITE,
                       "meta_convolver_ref": _lib.META_REF}[cfg.meta]
        c.feature_length = arch.FEATURE_LENGTH
        c.precision = _lib.PRECISIONS[precision]
        c.max_chunk_sites = max_chunk_sites
        handle = C.c_void_p()
        buf = (C.c_char * len(blob)).from_buffer_copy(blob)
        rc = self.lib.hello_moe_create(buf, len(blob), C.byref(c), index, C.byref(handle))
        if rc != 0:
            raise _lib.HelloMoEError("hello_moe_create failed (%d): %s" % (
                rc, self.lib.hello_moe_last_error(None).decode()))
        self.handle = handle
        self.workspace_cap = int(workspace_bytes)
        self._ws: Optional[torch.Tensor] = None

    def close(self):
        if getattr(self, "handle", None):
            self.lib.hello_moe_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- plumbing ---------------------------------------------------------------------------------------------
    def launch_count(self) -> int:
        return int(self.lib.hello_moe_launch_count(self.handle))

    def workspace_bytes(self, n_reads0: int, n_reads1: int, n_alleles: int, n_sites: int) -> int:
        return int(self.lib.hello_moe_workspace_bytes(self.handle, n_reads0, n_reads1, n_alleles, n_sites))

    def _workspace(self, need: int) -> torch.Tensor:
        want = min(max(need, 1 << 20), self.workspace_cap)
        if self._ws is None or self._ws.numel() < want:
            self._ws = None
            self._ws = torch.empty(want, dtype=torch.uint8, device=self.device)
        return self._ws

    def profile_enable(self, on: bool = True):
        self._check(self.lib.hello_moe_profile_enable(self.handle, int(on)), "hello_moe_profile_enable")

    def profile_collect(self):
        """(milliseconds spent in the read-convolver stage, number of bracketed regions) since the last collect."""
        ms, n = C.c_double(), C.c_int64()
        self._check(self.lib.hello_moe_profile_collect(self.handle, C.byref(ms), C.byref(n)),
                    "hello_moe_profile_collect")
        return ms.value, n.value

    def _check(self, rc: int, what: str):
        if rc != 0:
            raise _lib.HelloMoEError("%s failed (%d): %s" % (what, rc,
                                                            self.lib.hello_moe_last_error(self.handle).decode()))

    def alloc_result(self, b: DeviceBatch) -> BatchResult:
        dev, A, S, P = self.device, b.n_alleles, b.n_sites, b.n_pairs
        return BatchResult(
            logits=torch.empty((3, A), dtype=torch.float32, device=dev),
            meta=torch.empty((S, 3), dtype=torch.float32, device=dev),
            pair_prob=torch.empty((4, P), dtype=torch.float32, device=dev),
            pair_mix64=torch.empty((P,), dtype=torch.float64, device=dev),
            best_pair=torch.empty((S, 2), dtype=torch.int32, device=dev),
            best_prob=torch.empty((S,), dtype=torch.float32, device=dev),
            pair_off=b.pair_off_h,
            call_pair=torch.empty((S, 5, 2), dtype=torch.int32, device=dev),
            call_qual=torch.empty((S, 5), dtype=torch.float64, device=dev),
            best_expert=torch.empty((S,), dtype=torch.int32, device=dev))

    def result_from_views(self, b: DeviceBatch, views: Dict[str, torch.Tensor]) -> BatchResult:
        """A BatchResult whose tensors are caller-provided views (e.g. shard.SiteGatherer.result_views: one packed buffer
        per rank, so that the multi-GPU gather is a single collective)."""
        dev, A, S, P = self.device, b.n_alleles, b.n_sites, b.n_pairs
        want = {"logits": (3, A), "meta": (S, 3), "pair_prob": (4, P), "pair_mix64": (P,), "best_pair": (S, 2),
                "best_prob": (S,), "call_pair": (S, 5, 2), "call_qual": (S, 5), "best_expert": (S,)}
        for k, shp in want.items():
            t = views[k]
            if tuple(t.shape) != shp or t.device != dev or not t.is_contiguous():
                raise ValueError("result view %s has shape %s on %s, expected %s on %s" % (k, tuple(t.shape), t.device, shp, dev))
        return BatchResult(pair_off=b.pair_off_h, **{k: views[k] for k in want})

    def _abi_structs(self, b: DeviceBatch, out: BatchResult):
        n_tech = len(self.cfg.read_cin)
        if len(b.reads) < n_tech:
            raise ValueError("model needs %d technologies, batch has %d" % (n_tech, len(b.reads)))
        if self.cfg.meta == "meta_convolver_ref" and b.ref_onehot is None:
            raise ValueError("this model gates on the reference segment; reference_segments is required")
        hb = _lib.HelloBatch()
        hb.n_sites, hb.n_alleles = b.n_sites, b.n_alleles
        hb.input_layout = b.layout
        for t in range(n_tech):
            r = b.reads[t]
            ch = self.cfg.read_cin[t]
            if r.dtype != torch.uint8 or r.device != self.device or not r.is_contiguous():
                raise ValueError("reads must be contiguous uint8 tensors on %s" % self.device)
            want = (arch.FEATURE_LENGTH, ch) if b.layout == _lib.LAYOUT_RLC else (ch, arch.FEATURE_LENGTH)
            if tuple(r.shape[1:]) != want:
                raise ValueError("technology %d reads have shape %s, expected [R, %d, %d]" % (
                    t, tuple(r.shape), want[0], want[1]))
            hb.n_reads[t] = r.shape[0]
            hb.d_reads[t] = r.data_ptr()
            hb.d_allele_read_off[t] = b.allele_read_off_d[t].data_ptr()
            hb.h_allele_read_off[t] = b.allele_read_off_h[t].data_ptr()
            if b.allele_read_off_h[t].numel() != b.n_alleles + 1:
                raise ValueError("allele_read_off has the wrong length")
        hb.d_site_allele_off = b.site_allele_off_d.data_ptr()
        hb.h_site_allele_off = b.site_allele_off_h.data_ptr()
        hb.d_ref_onehot = b.ref_onehot.data_ptr() if b.ref_onehot is not None else None
        hb.d_allele_rank = b.allele_rank.data_ptr() if b.allele_rank is not None else None
        hb.d_pair_off = b.pair_off_d.data_ptr()
        hr = _lib.HelloResult()
        hr.d_logits, hr.d_meta = out.logits.data_ptr(), out.meta.data_ptr()
        hr.d_pair_prob, hr.d_pair_mix64 = out.pair_prob.data_ptr(), out.pair_mix64.data_ptr()
        hr.d_best_pair, hr.d_best_prob = out.best_pair.data_ptr(), out.best_prob.data_ptr()
        hr.d_call_pair, hr.d_call_qual = out.call_pair.data_ptr(), out.call_qual.data_ptr()
        hr.d_best_expert = out.best_expert.data_ptr()
        return hb, hr

    def run(self, b: DeviceBatch, out: Optional[BatchResult] = None,
            workspace: Optional[torch.Tensor] = None) -> BatchResult:
        """Enqueue the forward of a device-resident batch on the current stream (asynchronous)."""
        out = out or self.alloc_result(b)
        hb, hr = self._abi_structs(b, out)
        n_tech = len(self.cfg.read_cin)
        nr = [int(b.reads[t].shape[0]) if t < n_tech else 0 for t in range(2)]
        ws = workspace if workspace is not None else \
            self._workspace(self.workspace_bytes(nr[0], nr[1], b.n_alleles, b.n_sites))
        stream = torch.cuda.current_stream(self.device).cuda_stream
        with torch.cuda.device(self.device):
            rc = self.lib.hello_moe_forward(self.handle, C.byref(hb), C.byref(hr), ws.data_ptr(), ws.numel(),
                                            C.c_void_p(stream))
        self._check(rc, "hello_moe_forward")
        return out

    def run_range(self, b: DeviceBatch, out: BatchResult, s0: int, s1: int, workspace: torch.Tensor) -> None:
        """Enqueue the forward of sites [s0, s1) of a device batch (hello_moe_forward_range): all buffers describe the
        whole batch, only these sites are computed."""
        hb, hr = self._abi_structs(b, out)
        stream = torch.cuda.current_stream(self.device).cuda_stream
        with torch.cuda.device(self.device):
            rc = self.lib.hello_moe_forward_range(self.handle, C.byref(hb), C.byref(hr), s0, s1, b.n_pairs,
                                                  workspace.data_ptr(), workspace.numel(), C.c_void_p(stream))
        self._check(rc, "hello_moe_forward_range")

    def _scratch(self, in_bytes: int, out_bytes: int):
        """One pinned host + one device buffer for a small batch's inputs, and the same for its results.  Grown on demand,
        reused by every run_host call of this engine."""
        sc = getattr(self, "_sc", None)
        if sc is None or sc["h_in"].numel() < in_bytes or sc["h_out"].numel() < out_bytes:
            cap_in, cap_out = max(1 << 16, 2 * in_bytes), max(1 << 12, 2 * out_bytes)
            sc = {"h_in": torch.empty(cap_in, dtype=torch.uint8).pin_memory(),
                  "d_in": torch.empty(cap_in, dtype=torch.uint8, device=self.device),
                  "h_out": torch.empty(cap_out, dtype=torch.uint8).pin_memory(),
                  "d_out": torch.empty(cap_out, dtype=torch.uint8, device=self.device)}
            sc["np_in"], sc["np_out"] = sc["h_in"].numpy(), sc["h_out"].numpy()
            self._sc = sc
        return sc

    def run_host(self, reads: Sequence[np.ndarray], allele_read_off: Sequence[np.ndarray], site_allele_off: np.ndarray,
                 allele_rank: np.ndarray, ref_onehot: Optional[np.ndarray] = None):
        """Score a SMALL batch that lives in host memory (one site of the strict drop-in call, a few dozen sites of the
        scoring server) and return its results on the host: -> (raw result image uint8, {field: (offset, bytes, dtype, shape)}).

        The whole input travels as ONE image (reads, CSR arrays, allele ranks, pair offsets, reference one-hot; every piece
        16-byte aligned) in one host -> device copy from a pinned buffer, the results come back as one image in one device ->
        host copy, and the C ABI gets raw addresses into the two device images: no tensor is created for anything that only
        needs an address.  (A dozen small copies, allocations and views cost more than a site's kernels.)  Synchronous.
        reads[t]: uint8 [R_t, L, C_t]; allele_read_off[t]: int32 [A+1]; site_allele_off: int32 [S+1]; allele_rank: int32 [A]
        (tie-break order of the allele strings inside each site); ref_onehot: fp32 [S, L, 5] for the reference-gated model."""
        cfg, L = self.cfg, arch.FEATURE_LENGTH
        n_tech = len(cfg.read_cin)
        sao = np.asarray(site_allele_off, np.int64)
        S, A = sao.size - 1, int(sao[-1])
        napS = np.diff(sao)
        pair_off = np.zeros(S + 1, np.int64)
        np.cumsum(napS * (napS + 1) // 2, out=pair_off[1:])
        P = int(pair_off[-1])
        need_ref = cfg.meta == "meta_convolver_ref"
        if need_ref and ref_onehot is None:
            raise ValueError("this model gates on the reference segment; reference_segments is required")
        al = lambda x: (x + 15) & ~15
        off, o_reads, o_aro = 0, [], []
        for t in range(n_tech):
            r = reads[t]
            if r.dtype != np.uint8 or r.ndim != 3 or tuple(r.shape[1:]) != (L, cfg.read_cin[t]):
                raise ValueError("technology %d reads have shape %s %s, expected uint8 [R, %d, %d]" % (t, r.shape, r.dtype, L, cfg.read_cin[t]))
            o_reads.append(off); off = al(off + r.size)
        for t in range(n_tech):
            if np.asarray(allele_read_off[t]).size != A + 1:
                raise ValueError("allele_read_off has the wrong length")
            o_aro.append(off); off = al(off + 4 * (A + 1))
        o_sao = off; off = al(off + 4 * (S + 1))
        o_rank = off; off = al(off + 4 * A)
        o_po = off; off = al(off + 8 * (S + 1))
        o_ref = off
        if need_ref:
            off = al(off + 4 * S * L * 5)
        in_bytes = off
        fields, o = {}, 0
        for name, dt, shape in (("logits", np.float32, (3, A)), ("meta", np.float32, (S, 3)), ("pair_prob", np.float32, (4, P)),
                                ("pair_mix64", np.float64, (P,)), ("best_pair", np.int32, (S, 2)), ("best_prob", np.float32, (S,)),
                                ("call_pair", np.int32, (S, 5, 2)), ("call_qual", np.float64, (S, 5)), ("best_expert", np.int32, (S,))):
            nb = int(np.dtype(dt).itemsize)
            for d in shape:
                nb *= d
            fields[name] = (o, nb, dt, shape)
            o = al(o + nb)
        out_bytes = max(o, 16)
        sc = self._scratch(in_bytes, out_bytes)
        np_in = sc["np_in"]
        for t in range(n_tech):
            np_in[o_reads[t]:o_reads[t] + reads[t].size] = reads[t].reshape(-1)
            np_in[o_aro[t]:o_aro[t] + 4 * (A + 1)].view(np.int32)[:] = allele_read_off[t]
        np_in[o_sao:o_sao + 4 * (S + 1)].view(np.int32)[:] = sao
        np_in[o_rank:o_rank + 4 * A].view(np.int32)[:] = allele_rank
        np_in[o_po:o_po + 8 * (S + 1)].view(np.int64)[:] = pair_off
        if need_ref:
            np_in[o_ref:o_ref + 4 * S * L * 5].view(np.float32)[:] = np.asarray(ref_onehot, np.float32).reshape(-1)
        sc["d_in"][:in_bytes].copy_(sc["h_in"][:in_bytes], non_blocking=True)
        base_d, base_h, base_o = sc["d_in"].data_ptr(), sc["h_in"].data_ptr(), sc["d_out"].data_ptr()
        hb = _lib.HelloBatch()
        hb.n_sites, hb.n_alleles, hb.input_layout = S, A, _lib.LAYOUT_RLC
        for t in range(n_tech):
            hb.n_reads[t] = reads[t].shape[0]
            hb.d_reads[t] = base_d + o_reads[t]
            hb.d_allele_read_off[t], hb.h_allele_read_off[t] = base_d + o_aro[t], base_h + o_aro[t]
        hb.d_site_allele_off, hb.h_site_allele_off = base_d + o_sao, base_h + o_sao
        hb.d_ref_onehot = base_d + o_ref if need_ref else None
        hb.d_allele_rank, hb.d_pair_off = base_d + o_rank, base_d + o_po
        hr = _lib.HelloResult()
        hr.d_logits, hr.d_meta = base_o + fields["logits"][0], base_o + fields["meta"][0]
        hr.d_pair_prob, hr.d_pair_mix64 = base_o + fields["pair_prob"][0], base_o + fields["pair_mix64"][0]
        hr.d_best_pair, hr.d_best_prob = base_o + fields["best_pair"][0], base_o + fields["best_prob"][0]
        hr.d_call_pair, hr.d_call_qual = base_o + fields["call_pair"][0], base_o + fields["call_qual"][0]
        hr.d_best_expert = base_o + fields["best_expert"][0]
        nr = [int(reads[t].shape[0]) if t < n_tech else 0 for t in range(2)]
        ws = self._workspace(self.workspace_bytes(nr[0], nr[1], A, S))
        stream = torch.cuda.current_stream(self.device)
        with torch.cuda.device(self.device):
            rc = self.lib.hello_moe_forward(self.handle, C.byref(hb), C.byref(hr), ws.data_ptr(), ws.numel(),
                                            C.c_void_p(stream.cuda_stream))
        self._check(rc, "hello_moe_forward")
        sc["h_out"][:out_bytes].copy_(sc["d_out"][:out_bytes], non_blocking=True)
        stream.synchronize()
        return sc["np_out"][:out_bytes].copy(), fields, pair_off         # the caller keeps the results; the scratch is reused

    @staticmethod
    def _ranges(n_sites: int, chunk_sites: int):
        """Site ranges of a streamed call: the first ones are short (chunk/8, chunk/4, chunk/2, ...) because the first
        range's host -> device copy is the one transfer no computation hides."""
        size, s0 = max(1024, chunk_sites // 8), 0
        while s0 < n_sites:
            s1 = min(n_sites, s0 + min(size, chunk_sites))
            yield s0, s1
            s0, size = s1, size * 2

    def forward_host(self, hb: "HostBatch", chunk_sites: int = 65536, sync: bool = True,
                     fresh_result: bool = False) -> "HostResult":
        """End-to-end call on HOST buffers.  The device-side batch (read rows, CSR, results) is allocated once per
        HostBatch; every call streams the read rows host -> device in ranges of `chunk_sites` sites on a copy stream
        while the previous range computes (hello_moe_forward_range), then brings the per-site results back.

        With ``sync=True`` (default) the call returns when the results ARE in the returned host buffers.  With
        ``sync=False`` it returns as soon as everything is queued: the device -> host copies may still be in flight, so
        call ``result.wait()`` (or poll ``result.ready()``) before reading any field.
        The returned HostResult aliases pinned buffers cached on the HostBatch -- a second forward_host on the same batch
        overwrites them -- unless ``fresh_result=True`` asks for newly allocated ones."""
        st = hb.device_state(self)
        main = torch.cuda.current_stream(self.device)
        copy_s, comp_s = st["copy"], st["compute"]
        copy_s.wait_stream(main)
        comp_s.wait_stream(main)
        db, res, out = st["batch"], st["result"], hb.result_buffers(fresh=fresh_result)
        S = hb.n_sites
        with torch.cuda.stream(copy_s):                       # the CSR arrays first (a few MB), then the rows
            for dst, src in st["small"]:
                dst.copy_(src, non_blocking=True)
            small_done = torch.cuda.Event()
            small_done.record(copy_s)
        comp_s.wait_event(small_done)
        sao = hb.site_allele_off
        for s0, s1 in self._ranges(S, chunk_sites):
            a0, a1 = int(sao[s0]), int(sao[s1])
            ev = torch.cuda.Event()
            with torch.cuda.stream(copy_s):
                for t, r in enumerate(hb.reads):
                    aro = hb.allele_read_off[t]
                    r0, r1 = int(aro[a0]), int(aro[a1])
                    db.reads[t][r0:r1].copy_(r[r0:r1], non_blocking=True)
                if hb.ref_onehot is not None and db.ref_onehot is not None:
                    db.ref_onehot[s0:s1].copy_(hb.ref_onehot[s0:s1], non_blocking=True)
                ev.record(copy_s)
            with torch.cuda.stream(comp_s):
                comp_s.wait_event(ev)
                self.run_range(db, res, s0, s1, st["workspace"])
        with torch.cuda.stream(comp_s):
            for dst, src in zip(out.tensors(), res.tensors()):
                dst.copy_(src, non_blocking=True)
            done = torch.cuda.Event()
            done.record(comp_s)
        out.done_event = done
        main.wait_stream(comp_s)
        main.wait_stream(copy_s)
        if sync:
            done.synchronize()
        return out

    def forward_host_packed(self, hp: "HostPackedBatch", chunk_sites: int = 65536, sync: bool = True,
                            fresh_result: bool = False) -> "HostResult":
        """End-to-end call on HOST buffers holding ALIGNED READS instead of encoded rows: per range of `chunk_sites` sites
        the packed reads (bases, qualities, CIGARs, reference windows: ~370 bytes per row instead of 900) go host -> device
        on a copy stream, `hello_encode_reads` (include/hello_encode.h, the GPU form of the reference's
        computeFeaturesColoredSimple, c++/src/AlleleSearcherLiteFiltered.cpp:1031-1180) writes the [R, 150, C] rows of the
        range straight into the batch's device buffer on the compute stream, and `hello_moe_forward_range` scores the range.
        The 900-byte rows never exist in host memory or on PCIe.  Same result object and sync / fresh_result semantics as
        forward_host."""
        from . import encoder as E
        st = hp.device_state(self)
        main = torch.cuda.current_stream(self.device)
        copy_s, comp_s = st["copy"], st["compute"]
        copy_s.wait_stream(main)
        comp_s.wait_stream(main)
        db, res, out = st["batch"], st["result"], hp.result_buffers(fresh=fresh_result)
        S = hp.n_sites
        with torch.cuda.stream(copy_s):
            for dst, src in st["small"]:
                dst.copy_(src, non_blocking=True)
            small_done = torch.cuda.Event()
            small_done.record(copy_s)
        comp_s.wait_event(small_done)
        pk, dv = hp.packed_t, st["packed"]
        sao = hp.site_allele_off
        lib = E._load()
        L = arch.FEATURE_LENGTH
        for s0, s1 in self._ranges(S, chunk_sites):
            a0, a1 = int(sao[s0]), int(sao[s1])
            rb0, rb1 = int(hp.read_base[s0]), int(hp.read_base[s1])
            spans = {"bases": (int(hp.read_off[rb0]), int(hp.read_off[rb1])), "quals": (int(hp.read_off[rb0]), int(hp.read_off[rb1])),
                     "cigars": (int(hp.cigar_off[rb0]), int(hp.cigar_off[rb1])),
                     "reference": (int(hp.ref_off[s0]), int(hp.ref_off[s1]))}
            for f in ("read_off", "cigar_off"):
                spans[f] = (rb0, rb1 + 1)
            for f in ("ref_start", "mapq", "orientation", "hp"):
                spans[f] = (rb0, rb1)
            spans["ref_off"] = (s0, s1 + 1)
            for f in ("window_start", "assembly_start", "assembly_stop"):
                spans[f] = (s0, s1)
            ev = torch.cuda.Event()
            with torch.cuda.stream(copy_s):
                for f, (lo, hi) in spans.items():
                    dv[f][lo:hi].copy_(pk[f][lo:hi], non_blocking=True)
                for t in range(len(hp.row_read)):
                    aro = hp.allele_read_off[t]
                    r0, r1 = int(aro[a0]), int(aro[a1])
                    st["row_read"][t][r0:r1].copy_(hp.row_read[t][r0:r1], non_blocking=True)
                    st["row_site"][t][r0:r1].copy_(hp.row_site[t][r0:r1], non_blocking=True)
                if hp.ref_onehot is not None and db.ref_onehot is not None:
                    db.ref_onehot[s0:s1].copy_(hp.ref_onehot[s0:s1], non_blocking=True)
                ev.record(copy_s)
            with torch.cuda.stream(comp_s):
                comp_s.wait_event(ev)
                for t in range(len(hp.row_read)):
                    aro = hp.allele_read_off[t]
                    r0, r1 = int(aro[a0]), int(aro[a1])
                    if r1 > r0:
                        b = E.HelloEncodeBatch()
                        b.n_rows, b.feature_length, b.channels = r1 - r0, L, self.cfg.read_cin[t]
                        b.d_row_read = st["row_read"][t].data_ptr() + 4 * r0
                        b.d_row_site = st["row_site"][t].data_ptr() + 4 * r0
                        for f in ("read_off", "bases", "quals", "cigar_off", "cigars", "ref_start", "mapq", "orientation", "hp",
                                  "ref_off", "reference", "window_start", "assembly_start", "assembly_stop"):
                            setattr(b, "d_" + f, dv[f].data_ptr())
                        with torch.cuda.device(self.device):
                            rc = lib.hello_encode_reads(C.byref(b), db.reads[t].data_ptr() + r0 * L * self.cfg.read_cin[t],
                                                        C.c_void_p(comp_s.cuda_stream))
                        if rc != 0:
                            raise _lib.HelloMoEError("hello_encode_reads failed (%d): %s" % (rc, lib.hello_encode_last_error().decode()))
                self.run_range(db, res, s0, s1, st["workspace"])
        with torch.cuda.stream(comp_s):
            for dst, src in zip(out.tensors(), res.tensors()):
                dst.copy_(src, non_blocking=True)
            done = torch.cuda.Event()
            done.record(comp_s)
        out.done_event = done
        main.wait_stream(comp_s)
        main.wait_stream(copy_s)
        if sync:
            done.synchronize()
        return out

    def run_net(self, net: str, x: torch.Tensor, layout: int = _lib.LAYOUT_RLC) -> torch.Tensor:
        """Test hook: one sub-network on `x` (uint8 reads, or fp32 channel-last [n, L, C])."""
        nid = weights.NET_IDS[net]
        x = x.contiguous().to(self.device)
        n = x.shape[0]
        is_read = nid in (0, 1)
        lin = arch.FEATURE_LENGTH if is_read else x.shape[1]
        layers = self.cfg.networks()[net]
        co, lo = arch.net_out_shape(layers, lin)
        out = torch.empty((n, lo, co), dtype=torch.float32, device=self.device)
        ws = self._workspace(1 << 30)
        oc, ol = C.c_int32(), C.c_int32()
        stream = torch.cuda.current_stream(self.device).cuda_stream
        with torch.cuda.device(self.device):
            rc = self.lib.hello_moe_run_net(self.handle, nid, x.data_ptr(), n, lin, layout, out.data_ptr(),
                                            C.byref(oc), C.byref(ol), ws.data_ptr(), ws.numel(), C.c_void_p(stream))
        self._check(rc, "hello_moe_run_net")
        assert (oc.value, ol.value) == (co, lo), ((oc.value, ol.value), (co, lo))
        return out


    def readconv_debug(self, reads: torch.Tensor, phase: int = -1, layout: int = _lib.LAYOUT_RLC, tech: int = 0):
        """Test hook: the fused tensor-core read convolver alone.  Returns (features [R,36,64], dump or None) where the
        dump holds the post-activation values of layer phase `phase` as [ceil(R/3), 512, 64] (include/hello_moe.h)."""
        reads = reads.contiguous().to(self.device)
        n = reads.shape[0]
        out = torch.empty((n, 36, 64), dtype=torch.float32, device=self.device)
        dbg = None
        if phase >= 0:
            dbg = torch.zeros(((n + 2) // 3, 512, 64), dtype=torch.float32, device=self.device)
        stream = torch.cuda.current_stream(self.device).cuda_stream
        with torch.cuda.device(self.device):
            rc = self.lib.hello_moe_readconv_debug(self.handle, tech, reads.data_ptr(), n, layout, phase, out.data_ptr(),
                                                   dbg.data_ptr() if dbg is not None else None, C.c_void_p(stream))
        self._check(rc, "hello_moe_readconv_debug")
        return out, dbg


    def headconv_debug(self, net: str, x: torch.Tensor, phase: int = -1):
        """Test hook: one fused tensor-core head network (compressor / xattn on a combined input / meta_convolver /
        combiner on a concatenated [n,18,256] input) on fp32 channel-last items.  Returns (output, dump or None); the dump holds the post-activation values of layer
        phase `phase` as [groups, 256, 256] (include/hello_moe.h)."""
        nid = weights.NET_IDS[net]
        x = x.contiguous().float().to(self.device)
        n, lin, cin = x.shape
        co, lo = arch.net_out_shape(self.cfg.networks()[net], lin)
        out = torch.empty((n, lo, co), dtype=torch.float32, device=self.device)
        per = 12 if cin == 128 else 6
        dbg = None
        if phase >= 0:
            shape = (128, 512) if cin == 256 else (256, 256)        # combiner: its 512-channel intermediate
            dbg = torch.zeros(((n + per - 1) // per,) + shape, dtype=torch.float32, device=self.device)
        stream = torch.cuda.current_stream(self.device).cuda_stream
        with torch.cuda.device(self.device):
            rc = self.lib.hello_moe_headconv_debug(self.handle, nid, x.data_ptr(), n, phase, out.data_ptr(),
                                                   dbg.data_ptr() if dbg is not None else None, C.c_void_p(stream))
        self._check(rc, "hello_moe_headconv_debug")
        return out, dbg


@dataclass
class HostResult:
    """Per-site results in (pinned) host memory."""
    logits: torch.Tensor
    meta: torch.Tensor
    pair_prob: torch.Tensor
    pair_mix64: torch.Tensor
    best_pair: torch.Tensor
    best_prob: torch.Tensor
    call_pair: torch.Tensor
    call_qual: torch.Tensor
    best_expert: torch.Tensor
    done_event: Optional["torch.cuda.Event"] = None     # recorded after the last device -> host copy of the call

    def tensors(self):
        return (self.logits, self.meta, self.pair_prob, self.pair_mix64, self.best_pair, self.best_prob,
                self.call_pair, self.call_qual, self.best_expert)

    def nbytes(self) -> int:
        return sum(t.numel() * t.element_size() for t in self.tensors())

    def ready(self) -> bool:
        """True once the results of the forward_host call that produced this object are in host memory."""
        return self.done_event is None or self.done_event.query()

    def wait(self) -> "HostResult":
        """Block the host until the results are in host memory (needed after forward_host(..., sync=False))."""
        if self.done_event is not None:
            self.done_event.synchronize()
        return self


class HostBatch:
    """A ragged batch in HOST memory (pinned when possible) -- what a caller of the drop-in holds."""

    def __init__(self, reads: Sequence[torch.Tensor], layout: int, allele_read_off: Sequence[torch.Tensor],
                 site_allele_off: torch.Tensor, ref_onehot: Optional[torch.Tensor] = None, pin: bool = True):
        pin_ = (lambda t: t if t.is_pinned() else t.pin_memory()) if pin else (lambda t: t)
        self.reads = tuple(pin_(r.contiguous()) for r in reads)
        self.layout = layout
        self.allele_read_off = tuple(o.contiguous() for o in allele_read_off)
        self.site_allele_off = site_allele_off.contiguous()
        self.ref_onehot = pin_(ref_onehot.contiguous()) if ref_onehot is not None else None
        self.pair_off = pair_offsets(self.site_allele_off)
        self.pin = pin
        self._out: Optional[HostResult] = None

    @property
    def n_sites(self) -> int:
        return self.site_allele_off.numel() - 1

    def input_nbytes(self) -> int:
        n = sum(r.numel() for r in self.reads) + sum(o.numel() * 4 for o in self.allele_read_off)
        n += self.site_allele_off.numel() * 4
        if self.ref_onehot is not None:
            n += self.ref_onehot.numel() * 4
        return n

    def result_buffers(self, fresh: bool = False) -> HostResult:
        """Pinned host buffers for the per-site results.  The default returns the SAME cached buffers on every call
        (pinning is expensive); ``fresh=True`` allocates new ones the caller owns."""
        def make():
            A, S, P = int(self.site_allele_off[-1]), self.n_sites, int(self.pair_off[-1])
            mk = lambda shape, dt: torch.empty(shape, dtype=dt, pin_memory=self.pin)
            return HostResult(mk((3, A), torch.float32), mk((S, 3), torch.float32), mk((4, P), torch.float32),
                              mk((P,), torch.float64), mk((S, 2), torch.int32), mk((S,), torch.float32),
                              mk((S, 5, 2), torch.int32), mk((S, 5), torch.float64), mk((S,), torch.int32))
        if fresh:
            return make()
        if self._out is None:
            self._out = make()
        return self._out

    def device_state(self, engine: "MoEEngine"):
        """Device-side twin of this batch for engine.forward_host, allocated on first use: read-row buffers (filled
        range by range on every call), CSR arrays, result buffers, workspace, two streams."""
        st = getattr(self, "_dev", None)
        if st is not None and st["engine"] is engine:
            return st
        dev = engine.device
        pin = (lambda t: t if t.is_pinned() else t.pin_memory()) if self.pin else (lambda t: t)
        aro_h = tuple(pin(o) for o in self.allele_read_off)
        sao_h, po_h = pin(self.site_allele_off), pin(self.pair_off)
        aro_d = tuple(torch.empty_like(o, device=dev) for o in aro_h)
        sao_d, po_d = torch.empty_like(sao_h, device=dev), torch.empty_like(po_h, device=dev)
        need_ref = engine.cfg.meta == "meta_convolver_ref" and self.ref_onehot is not None
        batch = DeviceBatch(reads=tuple(torch.empty(r.shape, dtype=torch.uint8, device=dev) for r in self.reads),
                            layout=self.layout, allele_read_off_h=self.allele_read_off, allele_read_off_d=aro_d,
                            site_allele_off_h=self.site_allele_off, site_allele_off_d=sao_d, pair_off_h=self.pair_off,
                            pair_off_d=po_d,
                            ref_onehot=torch.empty(self.ref_onehot.shape, dtype=torch.float32, device=dev) if need_ref else None)
        small = [(d, h) for d, h in zip(aro_d, aro_h)] + [(sao_d, sao_h), (po_d, po_h)]
        ws = torch.empty(engine.workspace_cap, dtype=torch.uint8, device=dev)
        st = {"engine": engine, "batch": batch, "result": engine.alloc_result(batch), "small": small, "workspace": ws,
              "copy": torch.cuda.Stream(dev), "compute": torch.cuda.Stream(dev)}
        self._dev = st
        return st

    def device_chunk(self, s0: int, s1: int, device):
        """Upload sites [s0, s1) (asynchronously on the current stream) with chunk-local CSR offsets."""
        sao = self.site_allele_off
        a0, a1 = int(sao[s0]), int(sao[s1])
        reads, offs = [], []
        pin = (lambda t: t.pin_memory()) if self.pin else (lambda t: t)   # small; pinned so the upload stays async
        for t, r in enumerate(self.reads):
            aro = self.allele_read_off[t]
            r0, r1 = int(aro[a0]), int(aro[a1])
            reads.append(r[r0:r1])
            offs.append(pin(aro[a0:a1 + 1] - r0))
        ref = self.ref_onehot[s0:s1] if self.ref_onehot is not None else None
        db = DeviceBatch.from_host(reads, self.layout, offs, pin(sao[s0:s1 + 1] - a0), ref, device, pin=self.pin)
        return db, (a0, a1), (int(self.pair_off[s0]), int(self.pair_off[s1]))


class HostPackedBatch(HostBatch):
    """A ragged batch in HOST memory as ALIGNED READS (hello_b200.encoder.PackedReads) plus, per technology, the row plan
    that turns them into the network's rows (row -> read, row -> site; site -> allele -> supporting reads order, -1 = the
    all-zero row of an allele without support) and the usual CSR.  What a caller holds right after read sampling, before
    the reference would run computeFeaturesColoredSimple per allele."""

    def __init__(self, packed, row_read: Sequence, row_site: Sequence, allele_read_off: Sequence[torch.Tensor],
                 site_allele_off: torch.Tensor, ref_onehot: Optional[torch.Tensor] = None, pin: bool = True):
        pin_ = (lambda t: t if t.is_pinned() else t.pin_memory()) if pin else (lambda t: t)
        as_t = lambda a: a if torch.is_tensor(a) else torch.from_numpy(np.ascontiguousarray(a))
        self.packed = packed
        self.packed_t = {f: pin_(as_t(getattr(packed, f).view(np.int32) if getattr(packed, f).dtype == np.uint32
                                      else getattr(packed, f))) for f in packed.__dataclass_fields__ if f != "read_base"}
        self.read_base = np.asarray(packed.read_base, np.int64)
        self.read_off, self.cigar_off, self.ref_off = packed.read_off, packed.cigar_off, packed.ref_off
        self.row_read = tuple(pin_(as_t(r).to(torch.int32).contiguous()) for r in row_read)
        self.row_site = tuple(pin_(as_t(r).to(torch.int32).contiguous()) for r in row_site)
        self.reads = ()                                   # no encoded rows on the host
        self.layout = _lib.LAYOUT_RLC
        self.allele_read_off = tuple(o.contiguous() for o in allele_read_off)
        self.site_allele_off = site_allele_off.contiguous()
        self.ref_onehot = pin_(ref_onehot.contiguous()) if ref_onehot is not None else None
        self.pair_off = pair_offsets(self.site_allele_off)
        self.pin = pin
        self._out = None
        for t, rr in enumerate(self.row_read):
            if int(self.allele_read_off[t][-1]) != rr.numel() or self.row_site[t].numel() != rr.numel():
                raise ValueError("technology %d: the row plan does not match allele_read_off" % t)

    def input_nbytes(self) -> int:
        n = sum(t.numel() * t.element_size() for t in self.packed_t.values())
        n += sum(r.numel() * 4 for r in self.row_read) + sum(r.numel() * 4 for r in self.row_site)
        n += sum(o.numel() * 4 for o in self.allele_read_off) + self.site_allele_off.numel() * 4
        if self.ref_onehot is not None:
            n += self.ref_onehot.numel() * 4
        return n

    def device_state(self, engine: "MoEEngine"):
        st = getattr(self, "_dev", None)
        if st is not None and st["engine"] is engine:
            return st
        dev = engine.device
        pin = (lambda t: t if t.is_pinned() else t.pin_memory()) if self.pin else (lambda t: t)
        aro_h = tuple(pin(o) for o in self.allele_read_off)
        sao_h, po_h = pin(self.site_allele_off), pin(self.pair_off)
        aro_d = tuple(torch.empty_like(o, device=dev) for o in aro_h)
        sao_d, po_d = torch.empty_like(sao_h, device=dev), torch.empty_like(po_h, device=dev)
        need_ref = engine.cfg.meta == "meta_convolver_ref" and self.ref_onehot is not None
        L = arch.FEATURE_LENGTH
        rows = tuple(torch.empty((r.numel(), L, engine.cfg.read_cin[t]), dtype=torch.uint8, device=dev)
                     for t, r in enumerate(self.row_read))
        batch = DeviceBatch(reads=rows, layout=_lib.LAYOUT_RLC, allele_read_off_h=self.allele_read_off,
                            allele_read_off_d=aro_d, site_allele_off_h=self.site_allele_off, site_allele_off_d=sao_d,
                            pair_off_h=self.pair_off, pair_off_d=po_d,
                            ref_onehot=torch.empty(self.ref_onehot.shape, dtype=torch.float32, device=dev) if need_ref else None)
        small = [(d, h) for d, h in zip(aro_d, aro_h)] + [(sao_d, sao_h), (po_d, po_h)]
        st = {"engine": engine, "batch": batch, "result": engine.alloc_result(batch), "small": small,
              "workspace": torch.empty(engine.workspace_cap, dtype=torch.uint8, device=dev),
              "packed": {f: torch.empty_like(t, device=dev) for f, t in self.packed_t.items()},
              "row_read": tuple(torch.empty_like(r, device=dev) for r in self.row_read),
              "row_site": tuple(torch.empty_like(r, device=dev) for r in self.row_site),
              "copy": torch.cuda.Stream(dev), "compute": torch.cuda.Stream(dev)}
        self._dev = st
        return st


def _as_uint8(t: torch.Tensor) -> torch.Tensor:
    """The reference feeds bytes as floats (``tensors[0].float()``, :162); the kernels take the bytes."""
    if t.dtype == torch.uint8:
        return t
    u = t.to(torch.uint8)
    if not torch.equal(u.to(t.dtype), t):
        raise ValueError("read feature tensors must hold integers in [0, 255] (the C++ encoder's byte codes)")
    return u


class MoEAttentionB200:
    """Drop-in for ``MoEAttention``: same ``forward`` signature and return convention, computed on the GPU."""

    def __init__(self, cfg: arch.ModelConfig, params: Dict[str, torch.Tensor], device="cuda:0",
                 precision: str = "bf16x3", **engine_kwargs):
        self.cfg = cfg
        self.engine = MoEEngine(cfg, params, device, precision, **engine_kwargs)
        self.meta = object() if cfg.meta is not None else None   # wrapper tests `moeMerged.meta is not None`
        self.last_result: Optional[BatchResult] = None

    @classmethod
    def from_state_dict(cls, state_dict, softplus_nets=(), **kw) -> "MoEAttentionB200":
        """``softplus_nets``: the kinds of sub-network built with Softplus instead of ReLU, e.g. ("read_convolver", "xattn")
        -- a state dict does not say; ``load_wrapper`` reads it off the pickled modules.  Empty for every shipped model."""
        sd = weights.supported_state(state_dict, softplus_nets=softplus_nets)
        return cls(weights.cfg_from_state_dict(sd, softplus_nets=softplus_nets), sd, **kw)

    @classmethod
    def from_reference_module(cls, moe_attention, **kw) -> "MoEAttentionB200":
        """Build from a live reference ``MoEAttention`` (e.g. ``torch.load(...).moeMerged``)."""
        return cls.from_state_dict(moe_attention.state_dict(), **kw)

    def eval(self):
        return self

    def train(self, mode: bool = False):
        if mode:
            raise NotImplementedError("hello_b200 implements the inference forward only")
        return self

    def make_batch(self, tensors, numAllelesPerSite, numReadsPerAllele, reference_segments) -> DeviceBatch:
        n_tech = len(self.cfg.read_cin)
        reads, offs = [], []
        for t in range(n_tech):
            if tensors[t] is None or numReadsPerAllele[t] is None:
                raise ValueError("hybrid model called without technology %d tensors" % t)
            reads.append(_as_uint8(tensors[t]))
            offs.append(_csr(numReadsPerAllele[t]))
        napS = numAllelesPerSite.tolist() if torch.is_tensor(numAllelesPerSite) else list(numAllelesPerSite)
        sao = _csr(napS)
        A = int(sao[-1])
        for t in range(n_tech):
            if offs[t].numel() != A + 1 or int(offs[t][-1]) != reads[t].shape[0]:
                raise ValueError("numReadsPerAllele[%d] does not match tensors / numAllelesPerSite" % t)
        ref = reference_segments if self.cfg.meta == "meta_convolver_ref" else None
        if self.cfg.meta == "meta_convolver_ref" and ref is None:
            raise ValueError("reference_segments is required by this model")
        return DeviceBatch.from_host(reads, _lib.LAYOUT_RCL, offs, sao, ref, self.engine.device)

    def forward(self, tensors, numAllelesPerSite, numReadsPerAllele, reference_segments=None, *args, **kwargs):
        in_dev = tensors[0].device
        batch = self.make_batch(tensors, numAllelesPerSite, numReadsPerAllele, reference_segments)
        res = self.engine.run(batch)
        self.last_result = res
        A = batch.n_alleles
        logits = res.logits.to(in_dev)
        if self.cfg.returns_meta:
            return [logits[e].reshape(A, 1) for e in range(3)], res.meta.to(in_dev)
        head = 0 if self.cfg.xattn_present[0] else 2
        return logits[head].reshape(A, 1)

    __call__ = forward


CALL_NAMES = ("mixed", "expert0", "expert1", "expert2", "mean")


def allele_ranks(alleles_per_site: Sequence[Sequence[str]]) -> torch.Tensor:
    """Tie-break ranks for DeviceBatch.allele_rank: rank of every allele string inside its site's sorted order (the
    reference's ``sorted([(v, k)...], reverse=True)[0]`` falls back on the allele strings when values tie)."""
    out = []
    for names in alleles_per_site:
        order = sorted(range(len(names)), key=lambda i: names[i])
        rank = [0] * len(names)
        for r, i in enumerate(order):
            rank[i] = r
        out.extend(rank)
    return torch.tensor(out, dtype=torch.int32)


def final_calls(res: BatchResult, site: int, alleles: Sequence[str]) -> Dict[str, object]:
    """The final-call step for one site from the kernel's per-site records (prepareVcf.py:142-166, caller_calling.py:
    702-735): {"mixed" | "expert0..2" | "best" | "mean": ((allele_i, allele_j), QUAL)} and "choice" = np.argmax(meta)."""
    cp = res.call_pair[site].cpu().tolist()
    cq = res.call_qual[site].cpu().tolist()
    out = {name: ((alleles[cp[k][0]], alleles[cp[k][1]]), cq[k]) for k, name in enumerate(CALL_NAMES)}
    out["choice"] = int(res.best_expert[site])
    out["best"] = out["expert%d" % out["choice"]]
    return out


def feature_records(expert_prob, meta, pair_off, alleles_per_site: Sequence[Sequence[str]],
                    loci: Sequence[Tuple[str, int, int]]) -> List[Dict[str, object]]:
    """The ``.features`` list the caller pickles next to its VCF (caller_calling.py:746-754, 897-899) and that
    ``prepareVcf.vcfRecords`` (prepareVcf.py:110-175) reads back: one dict per site with ``chromosome``, ``position``,
    ``length`` (of the reference allele), ``meta`` (numpy float32 [3]) and ``expertPredictions`` (three dicts keyed by
    the (allele_i, allele_j) tuples, i <= j in site order, holding 0-d float32 tensors).

    expert_prob: [3, P] per-expert pair probabilities (rows 1..3 of BatchResult.pair_prob); meta: [S, 3];
    pair_off: [S+1]; loci: (chromosome, start, reference-allele length) per site."""
    ep = torch.as_tensor(expert_prob).detach().cpu().float()
    mt = torch.as_tensor(meta).detach().cpu().float().numpy()
    po = [int(x) for x in pair_off]
    if ep.dim() != 2 or ep.shape[0] != 3 or len(po) != len(alleles_per_site) + 1 or len(loci) != len(alleles_per_site):
        raise ValueError("feature_records: expert_prob must be [3, P] and pair_off / alleles / loci must describe the same sites")
    out = []
    for s, names in enumerate(alleles_per_site):
        n = len(names)
        if po[s + 1] - po[s] != n * (n + 1) // 2:
            raise ValueError("site %d: %d alleles do not match %d genotype pairs" % (s, n, po[s + 1] - po[s]))
        keys = [(names[i], names[j]) for i in range(n) for j in range(i, n)]
        chrom, start, length = loci[s]
        out.append({"chromosome": chrom, "position": int(start), "length": int(length), "meta": mt[s].copy(),
                    "expertPredictions": tuple({k: ep[e, po[s] + q] for q, k in enumerate(keys)} for e in range(3))})
    return out


def result_feature_records(res: "BatchResult", alleles_per_site, loci):
    """feature_records of a whole batch result (device or host)."""
    return feature_records(res.pair_prob[1:], res.meta, res.pair_off, alleles_per_site, loci)


def read_wrapper(path: str, reference_python: Optional[str] = None):
    """Unpickle a reference ``.wrapper.dnn`` on the CPU: -> (ModelConfig, weight-norm state dict, providePredictions).

    The files are whole-module pickles of ``MoEMergedWrapperAdvanced`` (python/create_model_wrapper.py:7-10,
    ``torch.save(wrapper, path)``), so unpickling needs the reference's own ``NNTools`` / ``MixtureOfExpertsAdvanced``
    importable -- a user of the reference has them on ``sys.path`` (it is how python/caller_calling.py:863 loads the same
    file); ``reference_python`` adds a directory for that.  ``NNTools`` must be imported first: it patches its layer
    classes onto ``torch.nn`` (python/NNTools.py:841-855) and the pickle refers to them there.  ``weights_only=False``
    because the file is a pickled module tree (with the legacy WeightNorm forward-pre-hooks), not a state dict."""
    import importlib
    import sys
    if reference_python and reference_python not in sys.path:
        sys.path.insert(0, reference_python)
    try:
        importlib.import_module("NNTools")
        importlib.import_module("MixtureOfExpertsAdvanced")
    except ImportError as exc:
        raise _lib.HelloMoEError(
            "load_wrapper: a .wrapper.dnn is a pickle of the reference's module tree; put the reference's python/ "
            "directory on sys.path (or pass reference_python=...) so that NNTools and MixtureOfExpertsAdvanced import: %s"
            % exc) from exc
    net = torch.load(path, map_location="cpu", weights_only=False)
    moe = getattr(net, "moeMerged", net)                     # the wrapper, or a bare MoEAttention
    if not hasattr(moe, "state_dict"):
        raise _lib.HelloMoEError("load_wrapper: %s does not hold a torch module" % path)
    # the activation and the kind of normalisation are not in a state dict: look at the modules
    kinds = {type(m).__name__ for m in moe.modules()}
    odd = sorted(kinds & {"LayerNorm", "LayerNormModule", "GroupNorm", "ELU", "LeakyReLU", "Tanh", "Sigmoid", "Dropout"})
    softplus_nets = set()
    for name, sub in moe.named_children():                   # the activation is chosen per sub-network (architecture module)
        acts = {type(m).__name__ for m in sub.modules()} & {"ReLU", "Softplus"}
        if acts == {"Softplus"}:
            if any(getattr(m, "beta", 1) != 1 or getattr(m, "threshold", 20) != 20 for m in sub.modules()
                   if type(m).__name__ == "Softplus"):
                odd.append("Softplus with non-default beta / threshold")
            softplus_nets.add(name.rstrip("0123456789"))
        elif len(acts) > 1:
            odd.append("ReLU and Softplus mixed inside " + name)
    if odd:
        raise _lib.HelloMoEError("load_wrapper: the model uses %s; this build covers the ReLU networks (weight-norm or "
                                 "BatchNorm1d) and the Softplus networks without normalisation layers that the reference's "
                                 "configurations build" % ", ".join(odd))
    sd = weights.supported_state(moe.state_dict(), softplus_nets=tuple(sorted(softplus_nets)))
    cfg = weights.cfg_from_state_dict(sd, softplus_nets=tuple(sorted(softplus_nets)))
    return cfg, sd, bool(getattr(net, "providePredictions", False))


def load_wrapper(path: str, device="cuda:0", precision: str = "bf16x3", reference_python: Optional[str] = None,
                 **engine_kwargs) -> "MoEMergedWrapperB200":
    """Drop-in for ``network = torch.load(path, map_location='cpu')`` (python/caller_calling.py:863-867): the same file,
    scored on the GPU.  ``.eval()`` and ``.providePredictions`` work as on the reference object."""
    cfg, sd, provide = read_wrapper(path, reference_python)
    moe = MoEAttentionB200(cfg, sd, device=device, precision=precision, **engine_kwargs)
    return MoEMergedWrapperB200(moe, providePredictions=provide)


class _SiteResult:
    """The results of one site scored through the wrapper: a host copy of the packed result image; the BatchResult fields
    (host tensors) are made on first access."""

    def __init__(self, raw: np.ndarray, fields, n_pairs: int):
        self._raw, self._fields = raw, fields
        self.pair_off = torch.tensor([0, n_pairs], dtype=torch.int64)

    def __getattr__(self, name):
        fields = self.__dict__.get("_fields", {})
        if name not in fields:
            raise AttributeError(name)
        o, nb, dt, shape = fields[name]
        t = torch.from_numpy(self._raw[o:o + nb].view(dt).reshape(shape))
        setattr(self, name, t)
        return t


class MoEMergedWrapperB200:
    """Drop-in for ``MoEMergedWrapperAdvanced``: ``network(featureDict, segment)`` for one site."""

    def __init__(self, moeMerged: MoEAttentionB200, providePredictions: bool = False):
        self.moeMerged = moeMerged
        self.providePredictions = providePredictions

    def eval(self):
        return self

    def forward(self, featureDict, segment):
        cfg = self.moeMerged.cfg
        eng = self.moeMerged.engine
        alleles = list(featureDict.keys())
        n, n_tech = len(alleles), len(cfg.read_cin)
        reads, aro = [], []
        for t in range(n_tech):
            parts = [featureDict[a][t] for a in alleles]
            if any(p is None for p in parts):
                raise ValueError("hybrid model called without technology %d tensors" % t)
            counts = [int(p.shape[0]) for p in parts]
            if min(counts) < 1:
                raise ValueError("every CSR slot must hold at least one row (reduceSlots requires it)")
            reads.append(_as_uint8(torch.cat(parts, dim=0)).contiguous().numpy())      # stays [r, L, C]: no transpose needed
            off = np.zeros(n + 1, np.int32)
            np.cumsum(counts, out=off[1:])
            aro.append(off)
        # tie-break of the reference's sort is on the allele strings (caller_calling.py:702-705)
        order = sorted(range(n), key=lambda i: alleles[i])
        rank = np.zeros(n, np.int32)
        for r, i in enumerate(order):
            rank[i] = r
        ref = segment.reshape(1, -1, 5).float().numpy() if cfg.meta == "meta_convolver_ref" else None
        raw, fields, pair_off = eng.run_host(reads, aro, np.array([0, n], np.int32), rank, ref)      # one H2D, one D2H
        P = int(pair_off[-1])
        self.moeMerged.last_result = _SiteResult(raw, fields, P)
        pp = torch.from_numpy(raw[fields["pair_prob"][0]:fields["pair_prob"][0] + 16 * P].view(np.float32).reshape(4, P))
        meta = torch.from_numpy(raw[fields["meta"][0]:fields["meta"][0] + 12].view(np.float32))
        keys = [(alleles[i], alleles[j]) for i in range(n) for j in range(i, n)]
        dicts = [{k: pp[row, q] for q, k in enumerate(keys)} for row in range(4)]
        bp = raw[fields["best_pair"][0]:fields["best_pair"][0] + 8].view(np.int32)
        self.last_call = (keys[self._pair_index(n, (int(bp[0]), int(bp[1])))],
                          float(raw[fields["best_prob"][0]:fields["best_prob"][0] + 4].view(np.float32)[0]))
        self.last_alleles = alleles
        if self.providePredictions:
            return tuple(dicts) + (meta,)
        return dicts[0]

    def final_calls(self) -> Dict[str, object]:
        """Calls of the site scored last, as prepareVcf.vcfRecords would make them from its .features record."""
        return final_calls(self.moeMerged.last_result, 0, self.last_alleles)

    @staticmethod
    def _pair_index(n: int, ij) -> int:
        i, j = ij
        return i * n - i * (i - 1) // 2 + (j - i)

    __call__ = forward
