"""Synthetic pileups in the byte format the reference's C++ encoder emits.

Channel code-book (c++/src/AlleleSearcherLiteFiltered.cpp:369-384, 971-1027, 1031-1180; Python spec
python/test_aligner.py:34-100):
  0 read base   A 250, G 180, T 100, C 30, gap 0        4 strand      + 70 / - 240
  1 ref base    same code                                5 position    allele span 240, elsewhere 70
  2 base qual   int(254*min(q,40)/40)                    6 haplotag    0 / 120 / 240 (HP models only)
  3 map qual    int(254*min(q,60)/60)
All channels are zero outside the read's span.  One-hot reference segment order is A,C,G,T,other
(python/caller_calling.py:53-67).

The distribution over sites is the one SURVEY.md 8(d) fixes: alleles/site ~ {1: .55, 2: .35, 3: .08, 4: .02},
reads/site ~ Poisson(coverage) with at least one read per allele.  Works on CPU and on CUDA (torch ops only);
it is plumbing for tests and the benchmark, not part of the product path.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Optional, Tuple

import numpy as np
import torch

from .arch import FEATURE_LENGTH

BASE_CODE = (250, 30, 180, 100)          # A, C, G, T  (index = one-hot column)
ALLELE_PROBS = (0.55, 0.35, 0.08, 0.02)


@dataclass
class Pileups:
    """A ragged batch of candidate sites (CSR).  reads[t]: uint8 [R_t, L, C_t] (the featureDict layout)."""
    reads: Tuple[torch.Tensor, ...]
    allele_read_off: Tuple[torch.Tensor, ...]   # int32 [A+1] per technology (host)
    site_allele_off: torch.Tensor               # int32 [S+1] (host)
    ref_onehot: torch.Tensor                    # fp32 [S, L, 5]

    @property
    def n_sites(self) -> int:
        return self.site_allele_off.numel() - 1

    @property
    def n_alleles(self) -> int:
        return int(self.site_allele_off[-1])

    def num_alleles_per_site(self):
        return torch.diff(self.site_allele_off).tolist()

    def num_reads_per_allele(self, tech: int = 0):
        return torch.diff(self.allele_read_off[tech]).tolist()

    def forward_args(self):
        """Arguments for the batched forward: (tensors [R,C,L], numAllelesPerSite, numReadsPerAllele, ref)."""
        tensors = tuple(r.transpose(1, 2) for r in self.reads)
        if len(tensors) == 1:
            tensors = (tensors[0], None)
            nrpa = (self.num_reads_per_allele(0), None)
        else:
            nrpa = (self.num_reads_per_allele(0), self.num_reads_per_allele(1))
        return tensors, self.num_alleles_per_site(), nrpa, self.ref_onehot

    def site_feature_dict(self, s: int, allele_names=None):
        """The per-site featureDict scoreSite builds (python/caller_calling.py:633-639)."""
        a0, a1 = int(self.site_allele_off[s]), int(self.site_allele_off[s + 1])
        names = allele_names or ["ACGT"[k % 4] * (1 + k // 4) for k in range(a1 - a0)]
        fd = {}
        for k, a in enumerate(range(a0, a1)):
            parts = []
            for t in range(len(self.reads)):
                r0, r1 = int(self.allele_read_off[t][a]), int(self.allele_read_off[t][a + 1])
                parts.append(self.reads[t][r0:r1].float().cpu())
            fd[names[k]] = (parts[0], parts[1] if len(parts) > 1 else None)
        return fd, self.ref_onehot[s:s + 1].cpu()


def _counts(n_sites: int, coverage, gen: torch.Generator, device, allow_empty_tech: bool):
    probs = torch.tensor(ALLELE_PROBS, device=device)
    n_alleles = torch.multinomial(probs, n_sites, replacement=True, generator=gen) + 1        # [S]
    if isinstance(coverage, (tuple, list)):
        lo, hi = coverage
        cov = torch.randint(lo, hi + 1, (n_sites,), generator=gen, device=device).float()
    else:
        cov = torch.full((n_sites,), float(coverage), device=device)
    n_reads = torch.poisson(cov, generator=gen).long()
    if allow_empty_tech:
        # a technology without support contributes one all-zero row per allele (AlleleSearcherLite.py:245-247)
        empty = torch.rand(n_sites, generator=gen, device=device) < 0.05
        n_reads = torch.where(empty, torch.zeros_like(n_reads), n_reads)
    else:
        empty = torch.zeros(n_sites, dtype=torch.bool, device=device)
    n_reads = torch.maximum(n_reads, n_alleles)
    return n_alleles, n_reads, empty


def _tech_reads(n_alleles, n_reads, empty, ref_idx, span_len, channels: int, gen, device):
    """Build uint8 [R, L, C] for one technology plus reads-per-allele counts."""
    L = FEATURE_LENGTH
    S = n_alleles.numel()
    site_of_read = torch.repeat_interleave(torch.arange(S, device=device), n_reads)
    R = site_of_read.numel()
    first_read = torch.cumsum(n_reads, 0) - n_reads
    q = torch.arange(R, device=device) - first_read[site_of_read]           # index of the read inside its site
    a_s = n_alleles[site_of_read]
    # first A_s reads seed one allele each; the rest favour allele 0 (the reference allele)
    u = torch.rand(R, generator=gen, device=device)
    pick = torch.where(u < 0.5, torch.zeros_like(a_s), (u * 2 - 1).mul(a_s).long().clamp_(max=3))
    pick = torch.minimum(pick, a_s - 1)
    allele_in_site = torch.where(q < a_s, q, pick)
    first_allele = torch.cumsum(n_alleles, 0) - n_alleles
    allele_of_read = first_allele[site_of_read] + allele_in_site
    A = int(n_alleles.sum())
    reads_per_allele = torch.bincount(allele_of_read, minlength=A)
    # rows must be grouped by allele: stable sort by global allele id
    order = torch.sort(allele_of_read, stable=True).indices
    site_of_read = site_of_read[order]

    pos = torch.arange(L, device=device)
    out = torch.empty((R, L, channels), dtype=torch.uint8, device=device)
    base_code = torch.tensor(BASE_CODE, dtype=torch.uint8, device=device)
    ref_codes = base_code[ref_idx]                                           # [S, L]
    read_ref = ref_codes[site_of_read]                                       # [R, L]
    mism = torch.rand((R, L), generator=gen, device=device) < 0.02
    rnd_base = base_code[torch.randint(0, 4, (R, L), generator=gen, device=device)]
    out[:, :, 0] = torch.where(mism, rnd_base, read_ref)
    out[:, :, 1] = read_ref
    qual = (torch.randn((R, L), generator=gen, device=device) * 6 + 32).clamp_(0, 40)
    out[:, :, 2] = (254 * qual / 40).to(torch.uint8)
    mapq = torch.randint(0, 61, (R, 1), generator=gen, device=device).float()
    out[:, :, 3] = (254 * mapq / 60).to(torch.uint8).expand(R, L)
    strand = torch.where(torch.rand((R, 1), generator=gen, device=device) < 0.5, 70, 240).to(torch.uint8)
    out[:, :, 4] = strand.expand(R, L)
    half = (span_len[site_of_read] // 2).unsqueeze(1)
    start = L // 2 - half
    in_span = (pos.unsqueeze(0) >= start) & (pos.unsqueeze(0) < start + span_len[site_of_read].unsqueeze(1))
    out[:, :, 5] = torch.where(in_span, 240, 70).to(torch.uint8)
    if channels == 7:
        hp = torch.randint(0, 3, (R, 1), generator=gen, device=device) * 120
        out[:, :, 6] = hp.to(torch.uint8).expand(R, L)
    # partial overlap: zero margins on either side
    left = torch.where(torch.rand(R, generator=gen, device=device) < 0.7, 0,
                       torch.randint(1, 61, (R,), generator=gen, device=device))
    right = torch.where(torch.rand(R, generator=gen, device=device) < 0.7, 0,
                        torch.randint(1, 61, (R,), generator=gen, device=device))
    covered = (pos.unsqueeze(0) >= left.unsqueeze(1)) & (pos.unsqueeze(0) < (L - right).unsqueeze(1))
    # technology with no support at this site: all-zero rows
    covered &= ~empty[site_of_read].unsqueeze(1)
    out *= covered.unsqueeze(2).to(torch.uint8)
    return out, reads_per_allele


def make_pileups(n_sites: int, coverage=30, channels=(6,), seed: int = 13, device="cpu",
                 uniform_bytes: bool = False) -> Pileups:
    """Generate `n_sites` synthetic candidate sites.  `channels` has one entry per technology (6 or 7);
    `coverage` is a mean or an inclusive (lo, hi) range drawn per site (BASELINE.json config 5)."""
    device = torch.device(device)
    gen = torch.Generator(device=device)
    gen.manual_seed(seed)                     # the reference seeds its own run with 13 (caller_calling.py:40-41)
    L = FEATURE_LENGTH
    ref_idx = torch.randint(0, 4, (n_sites, L), generator=gen, device=device)
    span_len = torch.where(torch.rand(n_sites, generator=gen, device=device) < 0.8, 1,
                           torch.randint(2, 11, (n_sites,), generator=gen, device=device))
    reads, offs = [], []
    n_alleles = None
    for t, ch in enumerate(channels):
        na, nr, empty = _counts(n_sites, coverage, gen, device, allow_empty_tech=(t == 1))
        if n_alleles is None:
            n_alleles = na
        nr = torch.maximum(nr, n_alleles)
        r, rpa = _tech_reads(n_alleles, nr, empty, ref_idx, span_len, ch, gen, device)
        if uniform_bytes:
            r = torch.randint(0, 256, r.shape, generator=gen, device=device, dtype=torch.uint8)
        reads.append(r)
        off = torch.zeros(rpa.numel() + 1, dtype=torch.int32)
        off[1:] = torch.cumsum(rpa, 0).to(torch.int32).cpu()
        offs.append(off)
    sao = torch.zeros(n_sites + 1, dtype=torch.int32)
    sao[1:] = torch.cumsum(n_alleles, 0).to(torch.int32).cpu()
    onehot = torch.nn.functional.one_hot(ref_idx, 5).float()
    return Pileups(tuple(reads), tuple(offs), sao, onehot)


def site_counts(n_sites: int, coverage=30, channels=(6,), seed: int = 13, device="cpu"):
    """(alleles per site, reads per site) of make_pileups(n_sites, coverage, channels, seed, device) WITHOUT building the read
    tensors: replays the head of the same random stream.  Used to cut a large synthetic dataset into balanced shards before
    any rank generates its own part (bench.py --partition balanced).  Single-technology workloads only: the counts of a
    second technology come after the first one's read bytes in the stream."""
    if len(channels) != 1:
        raise ValueError("site_counts replays the stream of a single-technology workload")
    device = torch.device(device)
    gen = torch.Generator(device=device)
    gen.manual_seed(seed)
    torch.randint(0, 4, (n_sites, FEATURE_LENGTH), generator=gen, device=device)              # ref_idx
    torch.rand(n_sites, generator=gen, device=device)                                         # span_len
    torch.randint(2, 11, (n_sites,), generator=gen, device=device)
    na, nr, _ = _counts(n_sites, coverage, gen, device, allow_empty_tech=False)
    return na, torch.maximum(nr, na)


def pair_offsets(site_allele_off: torch.Tensor) -> torch.Tensor:
    """int64 [S+1] prefix sum of A_s(A_s+1)/2 -- number of unordered genotype pairs per site."""
    n = torch.diff(site_allele_off.long())
    out = torch.zeros(n.numel() + 1, dtype=torch.int64)
    out[1:] = torch.cumsum(n * (n + 1) // 2, 0)
    return out


def packed_allele_csr(per_site_reads, seed: int = 13):
    """A synthetic allele structure over rows that are already in site order (one row per read): alleles per site from
    ALLELE_PROBS (clipped to the site's read count), every alternative allele gets max(1, n / (2A)) consecutive rows, the
    reference allele the rest.  -> (allele_read_off int32 [A+1], site_allele_off int32 [S+1])."""
    import numpy as np
    rng = np.random.default_rng(seed + 1)
    n = np.asarray(per_site_reads, np.int64)
    A = np.minimum(rng.choice(len(ALLELE_PROBS), size=n.size, p=np.asarray(ALLELE_PROBS) / np.sum(ALLELE_PROBS)) + 1, n)
    sao = np.zeros(n.size + 1, np.int64)
    np.cumsum(A, out=sao[1:])
    alt = np.maximum(1, n // (2 * A))
    site_of_allele = np.repeat(np.arange(n.size), A)
    k = np.arange(int(sao[-1])) - sao[site_of_allele]                   # allele index inside its site
    counts = np.where(k == 0, (n - (A - 1) * alt)[site_of_allele], alt[site_of_allele])
    aro = np.zeros(counts.size + 1, np.int64)
    np.cumsum(counts, out=aro[1:])
    assert (counts >= 1).all() and aro[-1] == n.sum()
    return torch.from_numpy(aro.astype(np.int32)), torch.from_numpy(sao.astype(np.int32))


def tile_packed_reads(packed, row_read, row_site, times: int):
    """`times` copies of a packed batch back to back (offsets rebased): a large synthetic batch from a small generated one."""
    import numpy as np
    from .encoder import PackedReads
    if times <= 1:
        return packed, row_read, row_site
    n_reads, n_sites = packed.ref_start.size, packed.window_start.size
    def offs(a):                                       # [n+1] offsets -> tiled offsets
        step = a[-1]
        return np.concatenate([a[:-1] + k * step for k in range(times)] + [np.array([times * step], a.dtype)])
    rep = lambda a: np.tile(a, times)
    out = PackedReads(offs(packed.read_off), rep(packed.bases), rep(packed.quals), offs(packed.cigar_off), rep(packed.cigars),
                      rep(packed.ref_start), rep(packed.mapq), rep(packed.orientation), rep(packed.hp), offs(packed.read_base),
                      offs(packed.ref_off), rep(packed.reference), rep(packed.window_start), rep(packed.assembly_start),
                      rep(packed.assembly_stop))
    rr = np.concatenate([np.where(row_read >= 0, row_read + k * n_reads, -1) for k in range(times)]).astype(np.int32)
    rs = np.concatenate([row_site + k * n_sites for k in range(times)]).astype(np.int32)
    return out, rr, rs


def make_packed_reads(n_sites: int, coverage: int = 30, read_len: int = 150, seed: int = 13, hp: bool = False):
    """Aligned reads for the GPU feature encoder, generated vectorised (numpy) straight in the packed layout of
    include/hello_encode.h: every read is `M a, I b, M c, D d, M e` with an insertion and / or a deletion in 15 % of the reads
    each; an absent operation is left out and the matches around it are merged, as in a BAM record (72 % of the reads are
    one match, 26 % have three operations, 2 % five), starting 0..119 bases before the 150-wide window centre.
    Returns (hello_b200.encoder.PackedReads, row_read int32, row_site int32) with one row per read in site order."""
    import numpy as np
    from .encoder import PackedReads
    rng = np.random.default_rng(seed)
    per_site = np.maximum(rng.poisson(coverage, n_sites), 1)
    R = int(per_site.sum())
    site_of = np.repeat(np.arange(n_sites, dtype=np.int32), per_site)
    ref_len = 450
    window_start = rng.integers(10_000, 200_000_000, n_sites).astype(np.int64)
    a0 = window_start + 225 + rng.integers(-10, 10, n_sites)
    a1 = a0 + rng.integers(1, 10, n_sites)
    start = (a0 + a1) // 2 - 75
    ins = np.where(rng.random(R) < 0.15, rng.integers(1, 8, R), 0)
    dele = np.where(rng.random(R) < 0.15, rng.integers(1, 8, R), 0)
    m1 = rng.integers(20, 60, R)
    m2 = rng.integers(20, 60, R)
    m3 = np.maximum(read_len - ins - m1 - m2, 1)
    lens = (m1 + ins + m2 + m3).astype(np.int64)
    ref_start = start[site_of] - rng.integers(0, 120, R)
    ref_start = np.maximum(ref_start, window_start[site_of] + 1)
    read_off = np.zeros(R + 1, np.int64)
    np.cumsum(lens, out=read_off[1:])
    total = int(read_off[-1])
    bases = np.frombuffer(b"ACGT", np.uint8)[rng.integers(0, 4, total)]
    quals = rng.integers(2, 42, total).astype(np.uint8)
    has_i, has_d = ins > 0, dele > 0
    # operation slots of a read, in order; a slot is kept when its operation exists (matches merged into the slot after a gap)
    ma = np.where(has_i, m1, 0)                                   # M before the insertion
    mc = np.where(has_d, np.where(has_i, m2, m1 + m2), 0)         # M between insertion and deletion
    me = np.where(has_d, m3, np.where(has_i, m2 + m3, m1 + m2 + m3))
    slots = np.stack([ma << 4, ins << 4 | 1, mc << 4, dele << 4 | 2, me << 4], axis=1).astype(np.uint32)
    keep = np.stack([has_i, has_i, has_d, has_d, np.ones(R, bool)], axis=1)
    cig = slots[keep]
    cigar_off = np.zeros(R + 1, np.int64)
    np.cumsum(keep.sum(axis=1), out=cigar_off[1:])
    read_base = np.zeros(n_sites + 1, np.int64)
    np.cumsum(per_site, out=read_base[1:])
    packed = PackedReads(read_off, bases, quals, cigar_off, cig, ref_start.astype(np.int64),
                         rng.integers(0, 61, R).astype(np.uint8), rng.choice(np.array([-1, 1], np.int8), R),
                         rng.integers(0, 3, R).astype(np.uint8) if hp else np.zeros(R, np.uint8), read_base,
                         np.arange(n_sites + 1, dtype=np.int64) * ref_len,
                         np.frombuffer(b"ACGT", np.uint8)[rng.integers(0, 4, n_sites * ref_len)],
                         window_start, a0.astype(np.int64), a1.astype(np.int64))
    return packed, np.arange(R, dtype=np.int32), site_of
