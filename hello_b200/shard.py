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


def site_costs(cfg: arch.ModelConfig, site_allele_off: np.ndarray, allele_read_offs: Sequence[np.ndarray]) -> np.ndarray:
    """Algorithmic FLOPs of every site: sum_t R_t*F_read_t + A*F_allele + F_site."""
    f_read, f_allele, f_site = arch.flops_model(cfg)
    sao = np.asarray(site_allele_off, dtype=np.int64)
    cost = np.diff(sao).astype(np.float64) * f_allele + f_site
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


def _all_gather_ragged(x: torch.Tensor, group=None) -> torch.Tensor:
    """Concatenate a per-rank tensor whose first dimension differs across ranks (counts exchanged first)."""
    world = dist.get_world_size(group)
    n = torch.tensor([x.shape[0]], dtype=torch.int64, device=x.device)
    counts = [torch.zeros_like(n) for _ in range(world)]
    dist.all_gather(counts, n, group=group)
    counts = [int(c.item()) for c in counts]
    cap = max(counts) if counts else 0
    if cap == 0:
        return x[:0]
    pad = torch.zeros((cap,) + tuple(x.shape[1:]), dtype=x.dtype, device=x.device)
    pad[:x.shape[0]] = x
    bufs = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(bufs, pad, group=group)
    return torch.cat([b[:c] for b, c in zip(bufs, counts)], dim=0)


@dataclass
class GatheredSites:
    best_pair: torch.Tensor    # int32 [S, 2]
    best_prob: torch.Tensor    # fp32 [S]
    meta: torch.Tensor         # fp32 [S, 3]
    pair_prob: torch.Tensor    # fp32 [P, 4]  (mixed, P_e0, P_e1, P_e2) per genotype pair, sites in global order
    pair_mix64: torch.Tensor   # fp64 [P]
    logits: torch.Tensor       # fp32 [A, 3]


def gather_site_results(best_pair: torch.Tensor, best_prob: torch.Tensor, meta: torch.Tensor,
                        pair_prob: torch.Tensor, pair_mix64: torch.Tensor, logits: torch.Tensor,
                        group=None) -> GatheredSites:
    """All-gather the per-site results of every rank (rank order == site order because shards are contiguous).
    Inputs are this rank's tensors: best_pair [S_r,2], best_prob [S_r], meta [S_r,3], pair_prob [4,P_r],
    pair_mix64 [P_r], logits [3,A_r]."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return GatheredSites(best_pair, best_prob, meta, pair_prob.t().contiguous(), pair_mix64,
                             logits.t().contiguous())
    return GatheredSites(
        best_pair=_all_gather_ragged(best_pair.contiguous(), group),
        best_prob=_all_gather_ragged(best_prob.contiguous(), group),
        meta=_all_gather_ragged(meta.contiguous(), group),
        pair_prob=_all_gather_ragged(pair_prob.t().contiguous(), group),
        pair_mix64=_all_gather_ragged(pair_mix64.contiguous(), group),
        logits=_all_gather_ragged(logits.t().contiguous(), group))


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

