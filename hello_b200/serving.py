"""Cross-process batching shim: many CPU caller processes, one GPU forward per batch of their sites.

The reference runs ``call.py --num_threads N``: a ``multiprocessing.Pool`` of N single-threaded workers
(python/call.py:111,217), each of which loads the network and calls ``network(featureDict, ref_segment)`` once per
site (python/caller_calling.py:651-652, 863-867).  One site per call leaves a GPU idle, so the drop-in for that
deployment is split in two:

  * ``ScoringServer`` -- one process owns the GPU engine.  It drains a request queue, concatenates the pending sites
    of all workers into one ragged batch (``BatchPlanner``), runs ``MoEEngine.run`` once and sends every worker its
    own sites' results back;
  * ``RemoteNetwork`` -- what a worker holds instead of the unpickled ``MoEMergedWrapperAdvanced``: same call
    signature, same ``.eval()`` / ``.providePredictions`` attributes, same return structure (dict keyed by allele
    tuples of 0-d tensors, or the 5-tuple ``(mixed, expert0, expert1, expert2, meta[3])``).

Results do not depend on which other sites share a batch (sites are independent; tests assert bit-equality with the
direct per-site call).  The batching logic is plain host code (``BatchPlanner``, ``serve_loop``) and is exercised on
the CPU with an injected scoring function; the product ``ScoringServer`` has no CPU mode.
"""
from __future__ import annotations

import multiprocessing as mp
import queue
import time
from dataclasses import dataclass
from typing import Callable, Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch

from . import arch

_STOP = "__stop__"


@dataclass
class SiteRequest:
    client: int
    seq: int
    alleles: List[str]
    reads: List[List[np.ndarray]]            # [technology][allele] -> uint8 [r, L, C]
    segment: Optional[np.ndarray]            # fp32 [1, L, 5] one-hot reference segment (gated hybrids) or None


def request_from_feature_dict(client: int, seq: int, featureDict, segment, n_tech: int) -> SiteRequest:
    """featureDict as caller_calling.py:633-639 builds it: {allele: (tensor [r,L,C], tensor [r',L,6] or None)}."""
    alleles = list(featureDict.keys())
    reads = []
    for t in range(n_tech):
        per = []
        for a in alleles:
            x = featureDict[a][t]
            if x is None:
                raise ValueError("hybrid model called without technology %d tensors" % t)
            x = x.detach().cpu() if torch.is_tensor(x) else torch.as_tensor(x)
            u = x.to(torch.uint8)
            if x.dtype != torch.uint8 and not torch.equal(u.to(x.dtype), x):
                raise ValueError("read feature tensors must hold integers in [0, 255] (the C++ encoder's byte codes)")
            per.append(u.contiguous().numpy())
        reads.append(per)
    seg = None if segment is None else np.ascontiguousarray(torch.as_tensor(segment).float().numpy())
    return SiteRequest(client, seq, alleles, reads, seg)


class BatchPlanner:
    """Concatenate pending per-site requests into the ragged batch layout of hello_moe_forward and split the per-site
    results back.  Pure host logic."""

    def __init__(self, n_tech: int):
        self.n_tech = n_tech
        self.reqs: List[SiteRequest] = []

    def add(self, req: SiteRequest):
        self.reqs.append(req)

    def __len__(self):
        return len(self.reqs)

    def build(self):
        """-> (reads per tech uint8 [R_t,L,C], allele_read_off per tech int32 [A+1], site_allele_off int32 [S+1],
        allele_rank int32 [A], ref_onehot fp32 [S,L,5] or None)."""
        reads, offs = [], []
        for t in range(self.n_tech):
            parts = [x for r in self.reqs for x in r.reads[t]]
            counts = [p.shape[0] for p in parts]
            if any(c < 1 for c in counts):
                raise ValueError("every allele needs at least one row per technology (reduceSlots requires it)")
            reads.append(torch.from_numpy(np.concatenate(parts, axis=0)))
            off = np.zeros(len(counts) + 1, np.int64)
            np.cumsum(counts, out=off[1:])
            offs.append(torch.from_numpy(off.astype(np.int32)))
        sao = np.zeros(len(self.reqs) + 1, np.int64)
        np.cumsum([len(r.alleles) for r in self.reqs], out=sao[1:])
        rank = []
        for r in self.reqs:               # the reference's sort falls back on the allele strings (caller_calling.py:702-705)
            order = sorted(range(len(r.alleles)), key=lambda i: r.alleles[i])
            rk = [0] * len(order)
            for pos, i in enumerate(order):
                rk[i] = pos
            rank.extend(rk)
        ref = None
        if all(r.segment is not None for r in self.reqs) and self.reqs:
            ref = torch.from_numpy(np.concatenate([r.segment.reshape(1, -1, 5) for r in self.reqs], axis=0))
        return (tuple(reads), tuple(offs), torch.from_numpy(sao.astype(np.int32)), torch.tensor(rank, dtype=torch.int32), ref)

    def split(self, pair_prob: np.ndarray, meta: np.ndarray, best_pair: np.ndarray, call_pair: np.ndarray,
              call_qual: np.ndarray, best_expert: np.ndarray):
        """pair_prob [4,P], meta [S,3], ... of the batch -> list of (request, per-site result dict)."""
        out, p0 = [], 0
        for s, r in enumerate(self.reqs):
            n = len(r.alleles)
            npairs = n * (n + 1) // 2
            out.append((r, {"pair_prob": pair_prob[:, p0:p0 + npairs].copy(), "meta": meta[s].copy(),
                            "best_pair": best_pair[s].copy(), "call_pair": call_pair[s].copy(),
                            "call_qual": call_qual[s].copy(), "best_expert": int(best_expert[s])}))
            p0 += npairs
        return out


def serve_loop(run_batch: Callable, n_tech: int, requests, responses: Sequence, max_sites: int = 4096,
               max_wait_s: float = 0.002, stats: Optional[dict] = None):
    """Drain `requests` (SiteRequest objects; the string _STOP ends the loop), score up to `max_sites` pending sites
    per call of `run_batch(reads, allele_read_off, site_allele_off, allele_rank, ref_onehot)` -> dict of numpy arrays
    (pair_prob, meta, best_pair, call_pair, call_qual, best_expert) and answer on responses[client]."""
    stop = False
    while not stop:
        first = requests.get()
        if isinstance(first, str) and first == _STOP:
            break
        plan = BatchPlanner(n_tech)
        plan.add(first)
        deadline = time.perf_counter() + max_wait_s
        while len(plan) < max_sites:
            try:
                nxt = requests.get(timeout=max(0.0, deadline - time.perf_counter()))
            except queue.Empty:
                break
            if isinstance(nxt, str) and nxt == _STOP:
                stop = True
                break
            plan.add(nxt)
        try:
            res = run_batch(*plan.build())
            for req, site in plan.split(res["pair_prob"], res["meta"], res["best_pair"], res["call_pair"], res["call_qual"],
                                        res["best_expert"]):
                responses[req.client].put((req.seq, site))
        except Exception as exc:                     # the worker re-raises: same failure path as a failing network call
            for req in plan.reqs:
                responses[req.client].put((req.seq, exc))
        if stats is not None:
            stats["batches"] = stats.get("batches", 0) + 1
            stats["sites"] = stats.get("sites", 0) + len(plan)


class RemoteNetwork:
    """Client-side stand-in for ``MoEMergedWrapperAdvanced``: ``network(featureDict, ref_segment)`` scores the site on the
    server's GPU.  Picklable; inherited by forked pool workers."""

    def __init__(self, requests, response, client: int, n_tech: int, has_meta_ref: bool):
        self._req, self._resp, self.client, self.n_tech, self._ref = requests, response, client, n_tech, has_meta_ref
        self.providePredictions = False
        self._seq = 0
        self.last_calls = None

    def eval(self):
        return self

    def __call__(self, featureDict, segment):
        self._seq += 1
        req = request_from_feature_dict(self.client, self._seq, featureDict, segment if self._ref else None, self.n_tech)
        self._req.put(req)
        seq, site = self._resp.get()
        if isinstance(site, Exception):
            raise site
        assert seq == self._seq, "response out of order"
        alleles = req.alleles
        n = len(alleles)
        keys = [(alleles[i], alleles[j]) for i in range(n) for j in range(i, n)]
        pp = torch.from_numpy(site["pair_prob"])
        dicts = [{k: pp[row, q] for q, k in enumerate(keys)} for row in range(4)]
        self.last_calls = site
        if self.providePredictions:
            return tuple(dicts) + (torch.from_numpy(site["meta"]),)
        return dicts[0]

    forward = __call__


def _gpu_server_main(cfg_name, params, device, precision, requests, responses, max_sites, max_wait_s, ready):
    from . import _lib, model
    cfg = arch.CONFIGS[cfg_name]
    net = model.MoEAttentionB200(cfg, params, device=device, precision=precision)

    def run_batch(reads, offs, sao, rank, ref):
        batch = model.DeviceBatch.from_host(reads, _lib.LAYOUT_RLC, offs, sao, ref if cfg.meta == "meta_convolver_ref" else None,
                                            net.engine.device, allele_rank=rank)
        r = net.engine.run(batch)
        torch.cuda.synchronize(net.engine.device)
        return {"pair_prob": r.pair_prob.cpu().numpy(), "meta": r.meta.cpu().numpy(), "best_pair": r.best_pair.cpu().numpy(),
                "call_pair": r.call_pair.cpu().numpy(), "call_qual": r.call_qual.cpu().numpy(),
                "best_expert": r.best_expert.cpu().numpy()}

    ready.set()
    serve_loop(run_batch, len(cfg.read_cin), requests, responses, max_sites, max_wait_s)


class ScoringServer:
    """Owns one GPU engine in its own (spawned) process.  ``client(i)`` gives worker i its ``RemoteNetwork``."""

    def __init__(self, cfg_name: str, params: Dict[str, torch.Tensor], n_clients: int, device="cuda:0",
                 precision: str = "bf16x3", max_sites: int = 4096, max_wait_s: float = 0.002):
        if not torch.cuda.is_available():
            from . import _lib
            raise _lib.HelloMoEError("ScoringServer needs a CUDA device (sm_100a); there is no CPU fallback")
        ctx = mp.get_context("spawn")
        self.cfg = arch.CONFIGS[cfg_name]
        self.requests = ctx.Queue()
        self.responses = [ctx.Queue() for _ in range(n_clients)]
        ready = ctx.Event()
        self.proc = ctx.Process(target=_gpu_server_main, daemon=True,
                                args=(cfg_name, {k: v.cpu() for k, v in params.items()}, device, precision, self.requests,
                                      self.responses, max_sites, max_wait_s, ready))
        self.proc.start()
        if not ready.wait(timeout=300):
            raise RuntimeError("scoring server did not come up")

    def client(self, i: int) -> RemoteNetwork:
        return RemoteNetwork(self.requests, self.responses[i], i, len(self.cfg.read_cin), self.cfg.meta == "meta_convolver_ref")

    def close(self):
        if self.proc.is_alive():
            self.requests.put(_STOP)
            self.proc.join(timeout=30)
            if self.proc.is_alive():
                self.proc.terminate()

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()
