"""CPU oracle for the HELLO MoE forward -- TEST INFRASTRUCTURE, NOT A PRODUCT PATH.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this file;
hello_b200/ never does.  It is a restatement, in plain torch-CPU fp32, of what the reference computes on the
path the CUDA library replaces.  The contractions themselves live in a third-party dependency of the
reference (PyTorch: Conv1d / Linear / MaxPool1d / AdaptiveAvgPool1d / softmax / sigmoid and
torch.nn.utils.weight_norm; the reference pins no version -- no requirements file, docker image
oddjobs/hello_deps, README.md:24), so the oracle calls the same torch.nn.functional ops on the CPU.

Parity pin: the reference publishes no golden vectors for this path.  The oracle is pinned
against outputs of the reference itself, produced in the build container by oracle/gen_golden.py (which imports
/root/reference/python unmodified) and committed under tests/golden/; tests/test_oracle_golden.py checks them.

Reference lines followed:
  reduce_slots        python/MixtureOfExpertsAdvanced.py:23-34
  run_net             python/NNTools.py:72-115 (conv+ReLU), 118-294 & 569-583 (residual blocks, no ReLU after
                      the add), 517-566 (terminus), 780-799 (weight norm)
  moe_forward         python/MixtureOfExpertsAdvanced.py:117-159, 161-252
  wrapper_forward     python/MixtureOfExpertsAdvanced.py:493-589
  call_genotype       python/caller_calling.py:702-705, 727-735
  remix_float64       python/prepareVcf.py:154-162
  final_calls         python/prepareVcf.py:59-62 (callAlleles: selection, QUAL), 142-166 (experts, best, mean);
                      pinned by tests/golden/final_calls.npz (made by oracle/gen_final_calls.py from the
                      reference's own vcfRecords)
"""
from __future__ import annotations

import itertools
import math
from typing import Dict, List, Optional, Sequence, Tuple

import torch
import torch.nn.functional as F

from hello_b200 import arch
from hello_b200.weights import folded


def reduce_slots(d: torch.Tensor, slots) -> torch.Tensor:
    """Segmented sum over contiguous row groups, computed as the reference does: cumulative sum over the whole
    batch, gathered at the slot ends and differenced."""
    slots = torch.as_tensor(slots, dtype=torch.long)
    csum = torch.cumsum(d, dim=0)
    ends = torch.cumsum(slots, dim=0) - 1
    picked = csum[ends]
    shifted = torch.cat((torch.zeros_like(d[:1]), picked[:-1]), dim=0)
    return picked - shifted


class FoldedNet:
    """One sub-network with weight-norm folded once (the reference refolds on every forward)."""

    def __init__(self, cfg: arch.ModelConfig, net: str, params):
        self.layers = []
        # the architecture modules' `activation` switch (NNTools.SingleConvLayer, python/NNTools.py:72-115): ReLU, or
        # torch.nn.Softplus() with its defaults for moe_attention_config_single_tech_old_equivalent_layer_norm.py
        self.act = {"relu": F.relu, "softplus": F.softplus}[cfg.activation(net)]
        for base, layer in cfg.keyed(net):          # addendum layers (XferLearning.py:131-160) simply follow
            if isinstance(layer, arch.Conv):
                self.layers.append(("conv", layer, folded(params, base + ".conv1d")))
            elif isinstance(layer, arch.MaxPool):
                self.layers.append(("pool", layer, None))
            elif isinstance(layer, arch.Res):
                wa = folded(params, base + ".ffNetwork.network.0.conv1d")
                wb = folded(params, base + ".ffNetwork.network.3.conv1d")
                ws = folded(params, base + ".shNetwork.network.0.conv1d") if layer.conv_shortcut else None
                self.layers.append(("res", layer, (wa, wb, ws)))
            elif isinstance(layer, arch.GapLinear):
                self.layers.append(("gap", layer, folded(params, arch.linear_key(base))))

    def _conv(self, x, c: arch.Conv, wb):
        y = F.conv1d(x, wb[0], wb[1], stride=c.stride, padding=c.pad)
        return self.act(y) if c.relu else y

    def __call__(self, x: torch.Tensor, trace: Optional[list] = None) -> torch.Tensor:
        """x: [n, C, L] fp32 (channel-major, as in the reference)."""
        for kind, layer, w in self.layers:
            if kind == "conv":
                x = self._conv(x, layer, w)
            elif kind == "pool":
                x = F.max_pool1d(x, layer.k, layer.stride, 0)
            elif kind == "res":
                wa, wb, ws = w
                ff = self._conv(self._conv(x, layer.conv_a, wa), layer.conv_b, wb)
                sh = self._conv(x, layer.conv_s, ws) if ws is not None else x
                x = ff + sh
            elif kind == "gap":
                x = F.adaptive_avg_pool1d(x, 1).view(x.shape[0], -1)
                x = F.linear(x, w[0], w[1])
            if trace is not None:
                trace.append(x)
        return x


class OracleModel:
    """MoEAttention restated (dead site-level compressor branch, :133-139, omitted: no shipped config reads it)."""

    def __init__(self, cfg: arch.ModelConfig, params: Dict[str, torch.Tensor]):
        self.cfg = cfg
        self.nets = {name: FoldedNet(cfg, name, params) for name in cfg.networks()}

    def _compress_and_predict(self, reduced, alleles_per_site: torch.Tensor, idx: int):
        c_allele = self.nets["compressor%d" % idx](reduced)
        c_site = reduce_slots(c_allele, alleles_per_site)
        logits = None
        if self.cfg.xattn_present[idx]:
            expanded = torch.repeat_interleave(c_site, alleles_per_site, dim=0)
            logits = self.nets["xattn%d" % idx](0 + 2 * c_allele + (-1) * expanded)
        return logits, c_site, c_allele

    @torch.no_grad()
    def forward(self, tensors, num_alleles_per_site, num_reads_per_allele, reference_segments=None):
        """Batched forward; same arguments and return convention as MoEAttention.forward."""
        cfg = self.cfg
        aps = torch.as_tensor(num_alleles_per_site, dtype=torch.long)
        conv0 = self.nets["read_convolver0"](tensors[0].float())
        red0 = reduce_slots(conv0, num_reads_per_allele[0])
        e0, s0, c0 = self._compress_and_predict(red0, aps, 0)
        if not cfg.hybrid:
            return e0
        conv1 = self.nets["read_convolver1"](tensors[1].float())
        red1 = reduce_slots(conv1, num_reads_per_allele[1])
        e1, s1, c1 = self._compress_and_predict(red1, aps, 1)
        e2, site_for_meta = None, None
        if cfg.xattn_present[2]:
            if cfg.legacy_sum:      # MoEMergedAdvanced.forward, useAdditive without ConvCombiners (MixtureOfExpertsAdvanced.py:408-436)
                c2 = c0 + c1
                s2 = reduce_slots(c2, aps)
            else:
                c2 = self.nets["combiner0"](torch.cat((c0, c1), dim=1))
                s2 = self.nets["combiner1"](torch.cat((s0, s1), dim=1))
            e2 = self.nets["xattn2"](0 + 2 * c2 + (-1) * torch.repeat_interleave(s2, aps, dim=0))
            site_for_meta = s2
        meta = None
        if cfg.meta == "meta_convolver":
            meta = torch.softmax(self.nets["meta"](site_for_meta), dim=-1)
        elif cfg.meta == "meta_convolver_ref":
            meta = torch.softmax(self.nets["meta"](reference_segments.float().transpose(1, 2)), dim=-1)
        if e0 is None and e1 is None:
            return e2
        if e2 is None:
            e2 = torch.zeros_like(e0)
        return [e0, e1, e2], meta

    @torch.no_grad()
    def read_features(self, tensor, tech: int = 0, trace: Optional[list] = None):
        return self.nets["read_convolver%d" % tech](tensor.float(), trace)


def pair_list(n: int) -> List[Tuple[int, int]]:
    """Genotype pairs in the reference's enumeration order: itertools.product with symmetric dedup = i<=j."""
    return [(i, j) for i in range(n) for j in range(i, n)]


def expert_probability(p: torch.Tensor, target: torch.Tensor) -> torch.Tensor:
    return torch.exp(torch.sum(torch.log(p * target + (1 - p) * (1 - target) + 1e-10)))


def site_posteriors(logits: Sequence[Optional[torch.Tensor]], meta: Optional[torch.Tensor]):
    """Per-site tail of the wrapper: sigmoid, pair enumeration, expert mixing (fp32).

    logits: three [A_s] tensors (None = expert absent -> probability 0, as torch.zeros_like in :536);
    meta: [3] or None (-> (1,0,0), :537-538).  Returns (mixed[P], experts[3][P], meta[3])."""
    n = next(l for l in logits if l is not None).numel()
    experts = [torch.sigmoid(l.reshape(-1)) if l is not None else torch.zeros(n) for l in logits]
    if meta is None:
        meta = torch.zeros(3)
        meta[0] = 1
    per_expert = []
    for e in range(3):
        vals = []
        for i, j in pair_list(n):
            t = torch.zeros(n)
            t[i] = 1
            t[j] = 1
            vals.append(expert_probability(experts[e], t))
        per_expert.append(torch.stack(vals))
    mixed = meta[0] * per_expert[0] + meta[1] * per_expert[1] + meta[2] * per_expert[2]
    return mixed, per_expert, meta


def wrapper_forward(model: OracleModel, feature_dict, segment, provide_predictions: bool = False):
    """MoEMergedWrapperAdvanced.forward restated: dict keyed by (allele_i, allele_j) string tuples."""
    alleles = list(feature_dict.keys())
    nr0 = [feature_dict[a][0].shape[0] for a in alleles]
    t0 = torch.cat([feature_dict[a][0].transpose(1, 2) for a in alleles], dim=0)
    if any(feature_dict[a][1] is None for a in alleles):
        t1, nr1 = None, [None] * len(alleles)
    else:
        nr1 = [feature_dict[a][1].shape[0] for a in alleles]
        t1 = torch.cat([feature_dict[a][1].transpose(1, 2) for a in alleles], dim=0)
    res = model.forward((t0, t1), [len(alleles)], (nr0, nr1), segment)
    if model.cfg.meta is not None:
        experts, meta = res
        logits, meta = [e.reshape(-1) for e in experts], meta[0]
    else:
        logits, meta = [res.reshape(-1), None, None], None
    mixed, per_expert, meta = site_posteriors(logits, meta)
    keys = [(alleles[i], alleles[j]) for i, j in pair_list(len(alleles))]
    mixed_d = {k: mixed[n] for n, k in enumerate(keys)}
    if not provide_predictions:
        return mixed_d
    return (mixed_d,) + tuple({k: pe[n] for n, k in enumerate(keys)} for pe in per_expert) + (meta,)


def call_genotype(pair_probs: Dict[Tuple[str, str], float]):
    """argmax exactly as caller_calling.py:702-705: sort (value, key) descending, take the first -- largest
    value, ties to the lexicographically greatest key.  Returns (key, value, qual)."""
    value, key = sorted(((float(v), k) for k, v in pair_probs.items()), reverse=True)[0]
    qual = -10 * math.log10(1 - min(value, 1 - 1e-8))
    return key, value, qual


def final_calls(expert_predictions, meta):
    """The final-call step of prepareVcf.vcfRecords (:142-166) without the VCF text: for each expert's pair
    probabilities, for the expert np.argmax(meta) picks ("best") and for the float64 re-mix ("mean"), the top
    allele pair and QUAL exactly as callAlleles computes them (:59-62).  expert_predictions: three dicts
    {(allele_i, allele_j): probability}; meta: three weights.  Returns {name: (top key, qual)} plus "choice"."""
    def call(likelihoods):
        likelihood, top = sorted([(float(v), k) for k, v in likelihoods.items()], reverse=True)[0]
        likelihood = min(float(likelihood), 1 - 1e-8)
        return top, -10 * math.log10(1 - likelihood)

    experts = [call(d) for d in expert_predictions]
    m = [float(x) for x in meta]
    choice = max(range(3), key=lambda i: (m[i], -i))                 # np.argmax: first maximum
    mean = {k: sum(float(expert_predictions[i][k]) * m[i] for i in range(3)) for k in expert_predictions[0]}
    return {"expert0": experts[0], "expert1": experts[1], "expert2": experts[2], "best": experts[choice],
            "mean": call(mean), "choice": choice}


def remix_float64(per_expert, meta) -> List[float]:
    """prepareVcf.py:154-162: sum_e float(P_e) * float(meta_e) in Python floats (float64)."""
    return [sum(float(per_expert[e][n]) * float(meta[e]) for e in range(3)) for n in range(len(per_expert[0]))]


def batched_posteriors(cfg: arch.ModelConfig, result, num_alleles_per_site, allele_rank=None):
    """Apply the wrapper tail to a batched MoEAttention-style result.  Returns per-site lists
    (mixed, experts, meta, best_pair, best_prob); tie-break rank defaults to the allele index."""
    if cfg.returns_meta:
        (e0, e1, e2), meta = result
        per_site_meta = True
    else:
        e0, e1, e2, meta, per_site_meta = result, None, None, None, False
    out, a0 = [], 0
    for s, n in enumerate(num_alleles_per_site):
        sl = slice(a0, a0 + n)
        logits = [e0[sl], e1[sl] if e1 is not None else None, e2[sl] if e2 is not None else None]
        mixed, per_expert, m = site_posteriors(logits, meta[s] if per_site_meta else None)
        rank = list(range(n)) if allele_rank is None else [int(r) for r in allele_rank[sl]]
        pairs = pair_list(n)
        best = max(range(len(pairs)), key=lambda q: (float(mixed[q]), rank[pairs[q][0]], rank[pairs[q][1]]))
        out.append((mixed, per_expert, m, pairs[best], float(mixed[best])))
        a0 += n
    return out
