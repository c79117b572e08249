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
import os
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

    def validate(self, req: SiteRequest) -> None:
        """Raise ValueError for a request that cannot be part of a batch (so that it is answered on its own instead of
        failing the whole batch it would have joined)."""
        n = len(req.alleles)
        if n < 1:
            raise ValueError("a site needs at least one allele")
        if len(req.reads) != self.n_tech or any(len(per) != n for per in req.reads):
            raise ValueError("featureDict does not hold %d technologies x %d alleles" % (self.n_tech, n))
        for t, per in enumerate(req.reads):
            for x in per:
                if x.ndim != 3 or x.shape[0] < 1:
                    raise ValueError("every allele needs at least one [r, L, C] row per technology (reduceSlots requires "
                                     "it; an allele without support carries one all-zero row)")
                if x.shape[1:] != per[0].shape[1:]:
                    raise ValueError("technology %d tensors of one site differ in shape" % t)

    def add(self, req: SiteRequest):
        self.validate(req)
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
               max_wait_s: float = 0.0003, stats: Optional[dict] = None, heartbeat=None, idle_tick_s: float = 0.5):
    """Drain `requests` (SiteRequest objects; the string _STOP ends the loop), score up to `max_sites` pending sites
    per call of `run_batch(reads, allele_read_off, site_allele_off, allele_rank, ref_onehot)` -> dict of numpy arrays
    (pair_prob, meta, best_pair, call_pair, call_qual, best_expert) and answer on responses[client].

    Batching policy: everything already queued is taken at once; the loop then waits at most `max_wait_s` for stragglers
    and never longer than it takes every client to have a request in the batch -- the workers call synchronously (one
    request in flight each, like the reference's pool workers), so once all of them are waiting nothing more can arrive.
    A forward of a few dozen sites costs the same as one of a single site (launch-bound), so requests that arrive while a
    batch is on the GPU simply form the next one; a long collection window only adds latency, and with synchronous
    callers latency IS throughput (sites/s = workers / round trip).

    A request that fails validation is answered with its exception on its own; if a whole batch fails, its sites are
    re-scored one by one so that only the offending client sees the exception.  `heartbeat` (a shared double) is stamped
    with time.time() at least every `idle_tick_s` seconds; clients use it to tell a busy server from a dead one."""
    def tick():
        if heartbeat is not None:
            heartbeat.value = time.time()

    def score(plan):
        t0 = time.perf_counter()
        built = plan.build()
        t1 = time.perf_counter()
        res = run_batch(*built)
        t2 = time.perf_counter()
        for req, site in plan.split(res["pair_prob"], res["meta"], res["best_pair"], res["call_pair"], res["call_qual"],
                                    res["best_expert"]):
            responses[req.client].put((req.seq, site))
        if stats is not None:                      # where a batch's time goes (seconds, summed over batches)
            t3 = time.perf_counter()
            for k, v in (("t_build", t1 - t0), ("t_forward", t2 - t1), ("t_respond", t3 - t2)):
                stats[k] = stats.get(k, 0.0) + v

    def admit(plan, item) -> bool:
        """item -> plan, or straight back to its client when it is malformed.  False on _STOP."""
        if isinstance(item, str) and item == _STOP:
            return False
        try:
            plan.add(item)
        except Exception as exc:
            responses[item.client].put((item.seq, exc))
        return True

    stop = False
    while not stop:
        tick()
        try:
            first = requests.get(timeout=idle_tick_s)
        except queue.Empty:
            continue
        t_first = time.perf_counter()
        plan = BatchPlanner(n_tech)
        if not admit(plan, first):
            break
        deadline = time.perf_counter() + max_wait_s
        full = min(max_sites, max(1, len(responses)))        # one outstanding request per client at most
        while len(plan) < max_sites:
            try:
                nxt = requests.get_nowait()                  # whatever queued up during the previous batch
            except queue.Empty:
                left = deadline - time.perf_counter()
                if left <= 0 or len(plan) >= full:
                    break
                try:
                    nxt = requests.get(timeout=left)
                except queue.Empty:
                    break
            if not admit(plan, nxt):
                stop = True
                break
        if len(plan) == 0:
            continue
        if stats is not None:
            stats["t_collect"] = stats.get("t_collect", 0.0) + time.perf_counter() - t_first
        tick()
        try:
            score(plan)
        except Exception:
            # one bad site must not poison up to max_sites others: score them one at a time, the worker of the failing
            # site re-raises (same failure path as a failing network call), everybody else gets a result
            for req in plan.reqs:
                single = BatchPlanner(n_tech)
                single.reqs.append(req)
                try:
                    score(single)
                except Exception as exc:
                    responses[req.client].put((req.seq, exc))
                tick()
        if stats is not None:
            stats["batches"] = stats.get("batches", 0) + 1
            stats["sites"] = stats.get("sites", 0) + len(plan)


class RemoteNetwork:
    """Client-side stand-in for ``MoEMergedWrapperAdvanced``: ``network(featureDict, ref_segment)`` scores the site on the
    server's GPU.  Picklable; inherited by forked pool workers."""

    def __init__(self, requests, response, client: int, n_tech: int, has_meta_ref: bool, server_pid: Optional[int] = None,
                 heartbeat=None, dead_after_s: float = 60.0):
        self._req, self._resp, self.client, self.n_tech, self._ref = requests, response, client, n_tech, has_meta_ref
        self._pid, self._hb, self._dead_after = server_pid, heartbeat, dead_after_s
        self.providePredictions = False
        self._seq = 0
        self.last_calls = None

    def eval(self):
        return self

    def _server_problem(self) -> Optional[str]:
        """Why the server cannot answer any more, or None.  Works from any process (forked pool workers are siblings of
        the server, not its parent): /proc tells a gone or zombie process, the heartbeat a hung one."""
        if self._pid is not None:
            try:
                with open("/proc/%d/stat" % self._pid) as f:
                    state = f.read().rsplit(")", 1)[1].split()[0]
                if state in ("Z", "X"):
                    return "the scoring server process (pid %d) has exited" % self._pid
            except FileNotFoundError:
                return "the scoring server process (pid %d) is gone" % self._pid
            except (OSError, IndexError):
                pass
        if self._hb is not None and self._hb.value > 0 and time.time() - self._hb.value > self._dead_after:
            return "the scoring server has not answered for %.0f s" % (time.time() - self._hb.value)
        return None

    def _await_response(self):
        while True:
            try:
                return self._resp.get(timeout=1.0)
            except queue.Empty:
                why = self._server_problem()
                if why:
                    raise RuntimeError("RemoteNetwork: %s; site not scored" % why)

    def __call__(self, featureDict, segment):
        self._seq += 1
        req = request_from_feature_dict(self.client, self._seq, featureDict, segment if self._ref else None, self.n_tech)
        self._req.put(req)
        seq, site = self._await_response()
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


def _gpu_server_main(cfg_name, params, device, precision, requests, responses, max_sites, max_wait_s, ready, heartbeat,
                     failure):
    from . import _lib, model
    cfg = arch.CONFIGS[cfg_name]
    try:
        net = model.MoEAttentionB200(cfg, params, device=device, precision=precision)
    except Exception as exc:                         # missing library / no GPU / bad weights: tell the parent why
        failure.put(repr(exc))
        raise

    def run_batch(reads, offs, sao, rank, ref):
        # one staging image per direction (MoEEngine.run_host): a batch of a few dozen sites costs one host -> device copy,
        # the kernels and one device -> host copy
        raw, fields, _ = net.engine.run_host([r.numpy() for r in reads], [o.numpy() for o in offs], sao.numpy(), rank.numpy(),
                                             ref.numpy() if (ref is not None and cfg.meta == "meta_convolver_ref") else None)
        view = lambda k: raw[fields[k][0]:fields[k][0] + fields[k][1]].view(fields[k][2]).reshape(fields[k][3])
        return {k: view(k) for k in ("pair_prob", "meta", "best_pair", "call_pair", "call_qual", "best_expert")}

    heartbeat.value = time.time()
    ready.set()
    stats = {} if os.environ.get("HELLO_SERVE_STATS") else None
    serve_loop(run_batch, len(cfg.read_cin), requests, responses, max_sites, max_wait_s, stats=stats, heartbeat=heartbeat)
    if stats:                                          # developer aid: where the server's time went, as JSON
        import json
        with open(os.environ["HELLO_SERVE_STATS"], "a") as f:
            f.write(json.dumps(stats) + "\n")


class ScoringServer:
    """Owns one GPU engine in its own (spawned) process.  ``client(i)`` gives worker i its ``RemoteNetwork``."""

    def __init__(self, cfg_name: str, params: Dict[str, torch.Tensor], n_clients: int, device="cuda:0",
                 precision: str = "bf16x3", max_sites: int = 4096, max_wait_s: float = 0.0003,
                 startup_timeout_s: float = 300.0):
        if not torch.cuda.is_available():
            from . import _lib
            raise _lib.HelloMoEError("ScoringServer needs a CUDA device (sm_100a); there is no CPU fallback")
        ctx = mp.get_context("spawn")
        self.cfg = arch.CONFIGS[cfg_name]
        self.requests = ctx.Queue()
        self.responses = [ctx.Queue() for _ in range(n_clients)]
        ready = ctx.Event()
        failure = ctx.Queue()
        self.heartbeat = ctx.Value("d", 0.0, lock=False)
        self.proc = ctx.Process(target=_gpu_server_main, daemon=True,
                                args=(cfg_name, {k: v.cpu() for k, v in params.items()}, device, precision, self.requests,
                                      self.responses, max_sites, max_wait_s, ready, self.heartbeat, failure))
        self.proc.start()
        t_end = time.time() + startup_timeout_s
        while not ready.wait(timeout=0.5):           # a child that died (no library, no GPU) is reported at once
            if not self.proc.is_alive():
                try:
                    why = failure.get(timeout=1.0)
                except queue.Empty:
                    why = "no message"
                raise RuntimeError("scoring server exited during start-up (exit code %s): %s" % (self.proc.exitcode, why))
            if time.time() > t_end:
                self.proc.terminate()
                raise RuntimeError("scoring server did not come up within %.0f s" % startup_timeout_s)

    def client(self, i: int) -> RemoteNetwork:
        return RemoteNetwork(self.requests, self.responses[i], i, len(self.cfg.read_cin), self.cfg.meta == "meta_convolver_ref",
                             server_pid=self.proc.pid, heartbeat=self.heartbeat)

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
