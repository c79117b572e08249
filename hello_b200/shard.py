"""Site sharding across GPUs and the one collective of the path: gathering per-site results.

Candidate sites are independent (no cross-site term anywhere in MoEAttention.forward; reduceSlots only groups
contiguous rows -- reference: python/MixtureOfExpertsAdvanced.py:23-34, 161-252), which is also how the
reference parallelises on the CPU: hotspot shards fanned out to a process pool (python/shardHotspots.py:78-137,
python/call.py:171-221).  Here a rank owns a contiguous range of sites chosen so that every rank gets about the
same amount of arithmetic; the forward runs without any communication; afterwards the fixed-size per-site
records (genotype call, its probability, expert weights) and the ragged genotype-pair probabilities are gathered
with torch.distributed (NCCL over NVLink on GPUs, gloo in the CPU tests).
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import List, Optional, Sequence, Tuple

import numpy as np
import torch
import torch.distributed as dist

from . import arch


def site_costs(cfg: arch.ModelConfig, site_allele_off: np.ndarray, allele_read_offs: Optional[Sequence[np.ndarray]] = None,
               reads_per_site: Optional[Sequence[np.ndarray]] = None) -> np.ndarray:
    """Algorithmic FLOPs of every site: sum_t R_t*F_read_t + A*F_allele + F_site.  The reads of a
    site are given either through the allele CSR (`allele_read_offs`, [A+1] per technology) or directly
    (`reads_per_site`, [S] per technology)."""
    f_read, f_allele, f_site = arch.flops_model(cfg)
    sao = np.asarray(site_allele_off, dtype=np.int64)
    cost = np.diff(sao).astype(np.float64) * f_allele + f_site
    if (allele_read_offs is None) == (reads_per_site is None):
        raise ValueError("give exactly one of allele_read_offs / reads_per_site")
    if reads_per_site is not None:
        for t, rps in enumerate(reads_per_site):
            cost += np.asarray(rps, dtype=np.float64) * f_read[t]
        return cost
    for t, aro in enumerate(allele_read_offs):
        aro = np.asarray(aro, dtype=np.int64)
        cost += (aro[sao[1:]] - aro[sao[:-1]]).astype(np.float64) * f_read[t]
    return cost


def balanced_ranges(cost: np.ndarray, world: int) -> List[Tuple[int, int]]:
    """Contiguous site ranges [s0, s1) per rank with near-equal total cost (boundaries on the cost prefix sum).
    Every site belongs to exactly one rank; ranks may be empty when there are fewer sites than ranks."""
    n = int(cost.shape[0])
    if world < 1:
        raise ValueError("world must be >= 1")
    prefix = np.concatenate(([0.0], np.cumsum(cost, dtype=np.float64)))
    total = prefix[-1]
    cuts = [0]
    for r in range(1, world):
        target = total * r / world
        s = int(np.searchsorted(prefix, target, side="left"))
        # choose the boundary nearest to the target, never moving backwards
        if s > 0 and abs(prefix[s - 1] - target) <= abs(prefix[min(s, n)] - target):
            s -= 1
        cuts.append(min(max(s, cuts[-1]), n))
    cuts.append(n)
    return [(cuts[r], cuts[r + 1]) for r in range(world)]


@dataclass
class SiteShard:
    """The slice of a ragged batch one rank owns (host-side CSR, offsets rebased to the shard)."""
    s0: int
    s1: int
    site_allele_off: torch.Tensor                 # int32 [S_r + 1]
    allele_read_off: Tuple[torch.Tensor, ...]     # int32 [A_r + 1] per technology
    read_range: Tuple[Tuple[int, int], ...]       # rows of reads[t] that belong to the shard
    allele_range: Tuple[int, int]


def take_shard(site_allele_off: torch.Tensor, allele_read_off: Sequence[torch.Tensor], s0: int, s1: int) -> SiteShard:
    sao = site_allele_off.to(torch.int64)
    a0, a1 = int(sao[s0]), int(sao[s1])
    offs, rr = [], []
    for aro in allele_read_off:
        aro = aro.to(torch.int64)
        r0, r1 = int(aro[a0]), int(aro[a1])
        offs.append((aro[a0:a1 + 1] - r0).to(torch.int32))
        rr.append((r0, r1))
    return SiteShard(s0, s1, (sao[s0:s1 + 1] - a0).to(torch.int32), tuple(offs), tuple(rr), (a0, a1))


# ------------------------------------------------------------------------------------------------ packed gather
# The per-site results of a rank live in ONE contiguous buffer (PackedSiteResults: the nine result tensors of
# hello_moe_forward are views into it), so gathering them is ONE collective: counts are exchanged once per batch shape,
# then every step does a single all_gather of `cap` bytes per rank -- on a side stream when asked, so that it overlaps the
# next step's forward.  (Round 1 did twelve list-form all_gathers per step, padded and re-concatenated, on the compute
# stream: 2 % of an 8-GPU step.)
RESULT_FIELDS = (
    # name, dtype, shape as a function of (S, A, P)
    ("logits", torch.float32, lambda S, A, P: (3, A)),
    ("meta", torch.float32, lambda S, A, P: (S, 3)),
    ("pair_prob", torch.float32, lambda S, A, P: (4, P)),
    ("pair_mix64", torch.float64, lambda S, A, P: (P,)),
    ("best_pair", torch.int32, lambda S, A, P: (S, 2)),
    ("best_prob", torch.float32, lambda S, A, P: (S,)),
    ("call_pair", torch.int32, lambda S, A, P: (S, 5, 2)),
    ("call_qual", torch.float64, lambda S, A, P: (S, 5)),
    ("best_expert", torch.int32, lambda S, A, P: (S,)),
)
_ALIGN = 256


def packed_layout(S: int, A: int, P: int):
    """-> ({name: (byte offset, shape, dtype)}, total bytes) of one rank's packed per-site results."""
    off, out = 0, {}
    for name, dt, shape in RESULT_FIELDS:
        shp = shape(S, A, P)
        n = int(np.prod(shp)) * torch.empty(0, dtype=dt).element_size()
        out[name] = (off, shp, dt)
        off = (off + n + _ALIGN - 1) // _ALIGN * _ALIGN
    return out, max(off, _ALIGN)


def packed_views(buf: torch.Tensor, S: int, A: int, P: int):
    """The nine result tensors as views into a uint8 buffer laid out by packed_layout."""
    layout, total = packed_layout(S, A, P)
    if buf.numel() < total:
        raise ValueError("packed result buffer too small")
    out = {}
    for name, (off, shp, dt) in layout.items():
        n = int(np.prod(shp)) * torch.empty(0, dtype=dt).element_size()
        out[name] = buf[off:off + n].view(dt).reshape(shp)
    return out


@dataclass
class GatheredSites:
    """Results of all ranks, sites in global order (rank order == site order because shards are contiguous)."""
    best_pair: torch.Tensor    # int32 [S, 2]
    best_prob: torch.Tensor    # fp32 [S]
    meta: torch.Tensor         # fp32 [S, 3]
    pair_prob: torch.Tensor    # fp32 [4, P]  (mixed, P_e0, P_e1, P_e2) per genotype pair
    pair_mix64: torch.Tensor   # fp64 [P]
    logits: torch.Tensor       # fp32 [3, A]
    call_pair: Optional[torch.Tensor] = None     # int32 [S, 5, 2]
    call_qual: Optional[torch.Tensor] = None     # fp64 [S, 5]
    best_expert: Optional[torch.Tensor] = None   # int32 [S]


class SiteGatherer:
    """One all_gather per step for the per-site results of every rank.

        g = SiteGatherer(device)                       # after init_process_group
        g.plan(S_r, A_r, P_r)                          # once per batch shape: exchanges the counts, allocates buffers
        res = g.result_views(slot)                     # dict of tensors to hand to hello_moe_forward (slot 0 / 1)
        g.gather_async(slot)                           # after the forward was queued on the current stream
        ...                                            # next step's forward into the other slot overlaps the gather
        g.wait(); all_ranks = g.concat(slot)           # when the gathered results are needed

    Works with NCCL (device tensors, side stream) and gloo (CPU tensors, synchronous) alike."""

    def __init__(self, device, group=None, slots: int = 2):
        self.device = torch.device(device)
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.slots = slots
        self.cuda = self.device.type == "cuda"
        self.stream = torch.cuda.Stream(self.device) if self.cuda else None
        self.counts = None
        self.collectives = 0

    def plan(self, S: int, A: int, P: int):
        mine = torch.tensor([S, A, P], dtype=torch.int64, device=self.device)
        if self.world > 1:
            allc = torch.empty(self.world * 3, dtype=torch.int64, device=self.device)
            dist.all_gather_into_tensor(allc, mine, group=self.group)
            self.collectives += 1
            self.counts = [tuple(int(x) for x in row) for row in allc.reshape(self.world, 3).cpu().tolist()]
        else:
            self.counts = [(S, A, P)]
        self.cap = max(packed_layout(*c)[1] for c in self.counts)
        self.local = [torch.zeros(self.cap, dtype=torch.uint8, device=self.device) for _ in range(self.slots)]
        self.gathered = [torch.empty(self.world * self.cap, dtype=torch.uint8, device=self.device)
                         for _ in range(self.slots)] if self.world > 1 else self.local
        self._done = [None] * self.slots
        return self

    def result_views(self, slot: int = 0):
        S, A, P = self.counts[self.rank]
        return packed_views(self.local[slot], S, A, P)

    def gather_async(self, slot: int = 0):
        """Queue the gather of `slot` behind the work already queued on the current stream; returns at once."""
        if self.world == 1:
            return
        if self.cuda:
            ready = torch.cuda.Event()
            ready.record(torch.cuda.current_stream(self.device))
            with torch.cuda.stream(self.stream):
                self.stream.wait_event(ready)
                dist.all_gather_into_tensor(self.gathered[slot], self.local[slot], group=self.group)
                done = torch.cuda.Event()
                done.record(self.stream)
            self._done[slot] = done
        else:
            dist.all_gather_into_tensor(self.gathered[slot], self.local[slot], group=self.group)
        self.collectives += 1

    def before_overwrite(self, slot: int):
        """Order the work queued next on the current stream (the forward that refills `slot`) behind the gather that is
        still reading it; no host wait."""
        if self._done[slot] is not None:
            torch.cuda.current_stream(self.device).wait_event(self._done[slot])

    def wait(self, slot: Optional[int] = None):
        """Make the current stream (and the host) wait for the queued gathers."""
        for k in range(self.slots) if slot is None else [slot]:
            if self._done[k] is not None:
                torch.cuda.current_stream(self.device).wait_event(self._done[k])
                self._done[k].synchronize()
                self._done[k] = None

    def rank_views(self, slot: int, rank: int):
        """Zero-copy views of the results `rank` contributed to the gathered buffer of `slot`."""
        buf = self.gathered[slot][rank * self.cap:(rank + 1) * self.cap] if self.world > 1 else self.local[slot]
        return packed_views(buf, *self.counts[rank])

    def concat(self, slot: int = 0) -> GatheredSites:
        """All ranks' results concatenated in site order (copies; the per-rank views are there for zero-copy use)."""
        v = [self.rank_views(slot, r) for r in range(self.world)]
        cat = lambda name, dim=0: torch.cat([x[name] for x in v], dim=dim)
        return GatheredSites(best_pair=cat("best_pair"), best_prob=cat("best_prob"), meta=cat("meta"),
                             pair_prob=cat("pair_prob", 1), pair_mix64=cat("pair_mix64"), logits=cat("logits", 1),
                             call_pair=cat("call_pair"), call_qual=cat("call_qual"), best_expert=cat("best_expert"))


def gather_site_results(best_pair: torch.Tensor, best_prob: torch.Tensor, meta: torch.Tensor,
                        pair_prob: torch.Tensor, pair_mix64: torch.Tensor, logits: torch.Tensor,
                        group=None) -> GatheredSites:
    """Convenience form for results that are NOT already packed: copies this rank's tensors (best_pair [S_r,2],
    best_prob [S_r], meta [S_r,3], pair_prob [4,P_r], pair_mix64 [P_r], logits [3,A_r]) into a packed buffer, then one
    count exchange + one all_gather.  Steady-state callers use SiteGatherer with result_views() directly."""
    S, A, P = int(best_pair.shape[0]), int(logits.shape[1]), int(pair_mix64.shape[0])
    g = SiteGatherer(best_pair.device, group=group, slots=1).plan(S, A, P)
    v = g.result_views(0)
    for name, t in (("best_pair", best_pair), ("best_prob", best_prob), ("meta", meta), ("pair_prob", pair_prob),
                    ("pair_mix64", pair_mix64), ("logits", logits)):
        v[name].copy_(t)
    g.gather_async(0)
    g.wait()
    out = g.concat(0)
    out.call_pair = out.call_qual = out.best_expert = None          # not supplied by this form
    return out


def parse_cpulist(text: str) -> List[int]:
    """'0-3,8,10-11' (the kernel's cpulist format) -> [0, 1, 2, 3, 8, 10, 11]."""
    cpus: List[int] = []
    for part in text.strip().split(","):
        if not part:
            continue
        lo, _, hi = part.partition("-")
        cpus.extend(range(int(lo), int(hi or lo) + 1))
    return cpus


def gpu_local_cpus(device_index: int, sysfs: str = "/sys/bus/pci/devices") -> List[int]:
    """CPU cores on the NUMA node the GPU hangs off (its PCI function's ``local_cpulist``); [] when unknown."""
    import os
    try:
        p = torch.cuda.get_device_properties(device_index)
        bdf = "%04x:%02x:%02x.0" % (p.pci_domain_id, p.pci_bus_id, p.pci_device_id)
        with open(os.path.join(sysfs, bdf, "local_cpulist")) as f:
            return parse_cpulist(f.read())
    except Exception:
        return []


def bind_rank_to_gpu_numa(device_index: int) -> List[int]:
    """One process per GPU: run this rank -- and, by first touch, place the pinned host buffers it allocates afterwards --
    on the socket its GPU is attached to, so host<->device copies do not cross the inter-socket link.  Keeps the current
    affinity when the topology is unknown or the intersection is empty.  Returns the cores now allowed."""
    import os
    allowed = set(os.sched_getaffinity(0))
    local = [c for c in gpu_local_cpus(device_index) if c in allowed]
    if local and len(local) < len(allowed):
        try:
            os.sched_setaffinity(0, local)
        except OSError:                   # a cpuset the kernel refuses: keep what we have
            pass
    return sorted(os.sched_getaffinity(0))

