"""Parameters of a HELLO MoE model: naming, deterministic init, weight-norm folding and the packed blob.

Parameter names are the reference's ``state_dict`` keys for ``MoEAttention``
(python/MixtureOfExpertsAdvanced.py:104-115; layers built by NNTools.Network, python/NNTools.py:633-657;
weight-normed layers hold ``weight_g``/``weight_v``/``bias`` under ``.conv1d`` / ``.linear``,
python/NNTools.py:780-799), so a real ``.wrapper.dnn`` state dict drops straight in.
"""
from __future__ import annotations

import struct
from typing import Dict, List, Tuple

import numpy as np
import torch

from . import arch

NET_IDS = {
    "read_convolver0": 0, "read_convolver1": 1, "compressor0": 2, "compressor1": 3,
    "xattn0": 4, "xattn1": 5, "xattn2": 6, "combiner0": 7, "combiner1": 8, "meta": 9,
}
N_NET_SLOTS = 10

KIND_CONV, KIND_MAXPOOL, KIND_RES, KIND_GAP_LINEAR = 0, 1, 2, 3
BLOB_MAGIC = b"HELLOB2\0"
BLOB_VERSION = 1
_HEADER_BYTES = 128
_REC_INTS = 32


def conv_keys(cfg: arch.ModelConfig, net: str) -> List[Tuple[str, Tuple[int, ...], str]]:
    """[(key prefix, v-shape, kind)] for every parametrised layer, in reference registration order."""
    out = []
    for base, layer in cfg.keyed(net):
        if isinstance(layer, arch.Conv):
            out.append((base + ".conv1d", (layer.cout, layer.cin, layer.k), "conv"))
        elif isinstance(layer, arch.Res):
            a, b, s = layer.conv_a, layer.conv_b, layer.conv_s
            out.append((base + ".ffNetwork.network.0.conv1d", (a.cout, a.cin, a.k), "conv"))
            out.append((base + ".ffNetwork.network.3.conv1d", (b.cout, b.cin, b.k), "conv"))
            if s is not None:
                out.append((base + ".shNetwork.network.0.conv1d", (s.cout, s.cin, s.k), "conv"))
        elif isinstance(layer, arch.GapLinear):
            out.append((arch.linear_key(base), (layer.cout, layer.cin), "linear"))
    return out


def param_shapes(cfg: arch.ModelConfig) -> Dict[str, Tuple[int, ...]]:
    """name -> shape for the whole model, mirroring MoEAttention.state_dict()."""
    shapes = {}
    for net in cfg.networks():
        for prefix, vshape, _ in conv_keys(cfg, net):
            shapes[prefix + ".bias"] = (vshape[0],)
            shapes[prefix + ".weight_g"] = (vshape[0],) + (1,) * (len(vshape) - 1)
            shapes[prefix + ".weight_v"] = vshape
    return shapes


def init_params(cfg: arch.ModelConfig, seed: int = 13) -> Dict[str, torch.Tensor]:
    """Random-init weights of the named architecture (BASELINE.json: the shipped blobs are git-lfs pointers).

    Machine-independent (numpy PCG64, float64 -> float32) so the same weights exist in this container, where the
    golden vectors are made with the reference, and on the GPU box.  Distribution mirrors torch's default
    Conv1d/Linear init (uniform, bound 1/sqrt(fan_in)); the weight-norm gain g is |v| scaled by U(0.75, 1.25) so
    that folding g*v/|v| is exercised with g != |v| as in a trained model.
    """
    rng = np.random.Generator(np.random.PCG64(seed))
    params = {}
    for net in cfg.networks():
        for prefix, vshape, _ in conv_keys(cfg, net):
            fan_in = int(np.prod(vshape[1:]))
            bound = 1.0 / np.sqrt(fan_in)
            v = ((rng.random(vshape) * 2.0 - 1.0) * bound).astype(np.float32)
            norm = np.sqrt((v.astype(np.float64) ** 2).reshape(vshape[0], -1).sum(axis=1))
            g = (norm * (0.75 + 0.5 * rng.random(vshape[0]))).astype(np.float32)
            b = ((rng.random(vshape[0]) * 2.0 - 1.0) * bound).astype(np.float32)
            params[prefix + ".weight_v"] = torch.from_numpy(v)
            params[prefix + ".weight_g"] = torch.from_numpy(g.reshape((vshape[0],) + (1,) * (len(vshape) - 1)))
            params[prefix + ".bias"] = torch.from_numpy(b)
    return params


def params_digest(params: Dict[str, torch.Tensor]) -> str:
    import hashlib
    h = hashlib.sha256()
    for k in sorted(params):
        h.update(k.encode())
        h.update(params[k].detach().cpu().contiguous().numpy().tobytes())
    return h.hexdigest()


def check_params(cfg: arch.ModelConfig, params: Dict[str, torch.Tensor]) -> None:
    shapes = param_shapes(cfg)
    missing = [k for k in shapes if k not in params]
    if missing:
        raise KeyError("missing parameters for %s: %s ..." % (cfg.name, missing[:4]))
    for k, shp in shapes.items():
        if tuple(params[k].shape) != shp:
            raise ValueError("parameter %s has shape %s, expected %s" % (k, tuple(params[k].shape), shp))


def fold_weight_norm(v: torch.Tensor, g: torch.Tensor) -> torch.Tensor:
    """w = g * v / |v|, norm over all dims but 0 -- torch.nn.utils.weight_norm(dim=0) as used by
    NNTools.WeightNormedConv1d/Linear (python/NNTools.py:783-785, 794-796).  Done once at load."""
    return torch._weight_norm(v.float(), g.float(), 0)


def folded(params: Dict[str, torch.Tensor], prefix: str) -> Tuple[torch.Tensor, torch.Tensor]:
    if prefix + ".weight" in params and prefix + ".weight_v" not in params:
        w = params[prefix + ".weight"].float()          # plain (non weight-normed) layer
    else:
        w = fold_weight_norm(params[prefix + ".weight_v"], params[prefix + ".weight_g"])
    return w, params[prefix + ".bias"].float()


def weight_norm_state(state_dict) -> Dict[str, torch.Tensor]:
    """A reference state dict reduced to what the packer reads: ``weight_g`` / ``weight_v`` / ``bias`` per layer.

    A materialised ``<layer>.weight`` next to its ``weight_g`` / ``weight_v`` (some torch versions keep the hook's last
    product in the state dict) is redundant and dropped.  Any other ``.weight`` -- a plain Conv1d/Linear, or the
    BatchNorm1d / LayerNorm layers of the configurations built without weight-norm (python/NNTools.py:27-45,84-104) --
    belongs to a model this build has no layer table for: refuse instead of silently losing the weights."""
    out, plain = {}, []
    for k, v in state_dict.items():
        if k.endswith(".weight"):
            if k[:-len(".weight")] + ".weight_v" in state_dict:
                continue
            plain.append(k)
        elif k.endswith((".running_mean", ".running_var", ".num_batches_tracked")):
            plain.append(k)
        else:
            out[k] = v
    if plain:
        raise ValueError("state dict holds layers without weight-norm (%s ...): BatchNorm / LayerNorm / plain-weight "
                         "configurations are not supported; every shipped model uses weight_norm = True" % plain[:3])
    return out


LEGACY_PREFIXES = ("readConv", "alleleConv", "expert", "siteConvCombiner.")     # (`meta.` exists in both wirings;
                                                                                 # "alleleConv" also covers alleleConvCombiner)
LEGACY_RENAME = {"readConv0": "read_convolver0", "readConv1": "read_convolver1", "alleleConv0": "compressor0",
                 "alleleConv1": "compressor1", "expert0": "xattn0", "expert1": "xattn1", "expert2": "xattn2", "meta": "meta",
                 "alleleConvCombiner": "combiner0", "siteConvCombiner": "combiner1"}


def legacy_state_to_attention(state_dict) -> Dict[str, torch.Tensor]:
    """State dict of the legacy wiring ``MoEMergedAdvanced`` (python/MixtureOfExpertsAdvanced.py:255-484, built by
    ``createMoEFullMergedAdvancedModel`` :614-654 from MoEReadConvolverDeeper / ExpertAlleleConvolverDeeper /
    ExpertGraphConvolverDeeper / MetaCombinerDeeper with ``useAdditive``) -> the ``MoEAttention`` names this package uses.

    The legacy sub-networks have the layers of the live ones and register their parameters in the same order, so the k-th
    tensor of ``readConv<t> / alleleConv<t> / expert<e> / meta`` is the k-th tensor of ``read_convolver<t> / compressor<t> /
    xattn<e> / meta`` (only the Sequential slot numbers differ: the live xattn / meta have parameter-free front-end modules).
    One technology: the legacy model computes exactly MoEAttention's function (expert input ``a - (s - a)``, :374-379) and maps
    onto ``single_tech`` / ``single_tech_hp``.  Two technologies (three experts + meta): the hybrid allele feature is the SUM of
    the two technologies' and the hybrid site frame its per-site sum (:408-436) -- ``legacy_hybrid_additive``.  With BOTH
    ``alleleConvCombiner`` and ``siteConvCombiner`` (``ConvCombiner`` = concatenate channels, then a network, :37-44; the
    reference's ``ConvCombinerResNetDeeper`` has the layers of ``architectures/conv_combiner.py``) the hybrid allele feature is
    ``alleleConvCombiner(a0, a1)``, the hybrid site frame ``siteConvCombiner(s0, s1)`` and meta reads that frame (:408-459):
    exactly MoEAttention's three-expert wiring -- ``hybrid_full``.  The factory builds ``meta`` without weight-norm
    (``make_network(configDict, "meta")``, :622), i.e. with BatchNorm1d even in a weight-norm model, and the legacy combiner
    modules have no weight-norm switch at all: such sub-networks are folded (``_fold_batchnorm_net``).  A legacy hybrid with
    only ONE of the two combiners, the concatenating form (``useAdditive=False``: experts twice as wide, no layer table here)
    and separate meta read convolvers (``readConv*Meta``) are refused.
    A state dict that is not legacy is returned unchanged."""
    keys = list(state_dict.keys())
    if not keys or not any(k.startswith(LEGACY_PREFIXES) for k in keys):
        return dict(state_dict)
    rename = LEGACY_RENAME
    groups = {}
    for k in keys:
        groups.setdefault(k.split(".", 1)[0], []).append(k)
    extra = sorted(set(groups) - set(rename))
    n_comb = ("alleleConvCombiner" in groups) + ("siteConvCombiner" in groups)
    if extra or n_comb == 1:
        raise ValueError("legacy MoEMergedAdvanced wiring with %s is not supported (separate meta read convolvers, or only one "
                         "of alleleConvCombiner / siteConvCombiner); supported: one technology, or two technologies with "
                         "additive features, summed or combined by both ConvCombiners"
                         % (extra or sorted(g for g in groups if g.endswith("Combiner"))))
    want = {rename[g] for g in groups}
    for cfg in arch.CONFIGS.values():
        # two technologies: summed hybrid features (legacy_sum) without combiners, MoEAttention's own wiring with them
        if set(cfg.networks()) != want or cfg.addendum or cfg.width != 1 or cfg.softplus_nets or \
                (cfg.hybrid and cfg.legacy_sum != (n_comb == 0)):
            continue
        out, ok = {}, True
        shapes = param_shapes(cfg)
        for old, ks in groups.items():
            new = rename[old]
            ours = conv_keys(cfg, new)
            if any(k.endswith(".running_mean") for k in ks):
                ok = _emit_folded(out, _fold_batchnorm_net(ks, state_dict), ours)
            else:
                theirs = [k for k in ks if not (k.endswith(".weight") and k[:-len(".weight")] + ".weight_v" in state_dict)]
                names = []
                for prefix, _, _ in ours:
                    names += [prefix + ".bias", prefix + ".weight_g", prefix + ".weight_v"]
                ok = len(theirs) == len(names) and all(tuple(state_dict[a].shape) == shapes[b] and
                                                       a.rsplit(".", 1)[1] == b.rsplit(".", 1)[1] for a, b in zip(theirs, names))
                if ok:
                    out.update({b: state_dict[a] for a, b in zip(theirs, names)})
            if not ok:
                break
        if ok:
            return {k: out[k] for k in shapes}              # the live model's registration order
    raise ValueError("legacy state dict does not match any supported configuration")


BN_EPS = 1e-5          # torch.nn.BatchNorm1d default, what NNTools builds (python/NNTools.py:84-104)


def is_batchnorm_state(state_dict) -> bool:
    return any(k.endswith(".running_mean") for k in state_dict)


def _fold_batchnorm_net(keys, state_dict, eps: float = BN_EPS):
    """[(w', b')] of one sub-network's convolution / linear layers in registration order, batch-norms folded:
      conv -> BN (same Sequential, next slot):   w' = w * s[o],  b' = (b - mean) * s + beta,   s = gamma / sqrt(var + eps)
      BN -> linear (pooled head, AvgPool -> Flatten -> BN -> Linear):  w' = w * s[i],  b' = b + w @ (beta - mean * s)
    A convolution without a following BN (the 1x1 shortcut of a stride-2 residual block) is taken as it is."""
    def parent_slot(stem):
        parent, slot = stem.rsplit(".", 1)
        return parent, int(slot) if slot.isdigit() else None

    stems = []
    for k in keys:
        st = k.rsplit(".", 1)[0]
        if not stems or stems[-1] != st:
            stems.append(st)
    layers = []                 # [w, b, stem, folded?] in order
    pre_bn = None               # (scale, shift, stem) waiting for the linear of the pooled head
    for st in stems:
        g = lambda name: state_dict.get(st + "." + name)
        if g("running_mean") is not None:
            gamma, beta, mean, var = (g("weight").double(), g("bias").double(), g("running_mean").double(),
                                      g("running_var").double())
            sc = gamma / torch.sqrt(var + eps)
            par, slot = parent_slot(st)
            if layers and not layers[-1][3] and parent_slot(layers[-1][2]) == (par, slot - 1 if slot is not None else None) \
                    and layers[-1][0].shape[0] == sc.numel():
                w, b = layers[-1][0], layers[-1][1]
                layers[-1][0] = w * sc.reshape((-1,) + (1,) * (w.dim() - 1))
                layers[-1][1] = (b - mean) * sc + beta
                layers[-1][3] = True
            else:
                pre_bn = (sc, beta - mean * sc, st)
        elif g("weight") is not None and g("weight").dim() >= 2:
            w, b = g("weight").double(), g("bias").double()
            if pre_bn is not None:
                sc, sh, bst = pre_bn
                if w.dim() != 2 or w.shape[1] != sc.numel() or parent_slot(bst)[0] != parent_slot(st)[0]:
                    raise ValueError("BatchNorm %s is followed by %s: only conv -> BN and BN -> linear are folded" % (bst, st))
                b = b + w @ sh
                w = w * sc.reshape(1, -1)
                pre_bn = None
            layers.append([w, b, st, False])
        else:
            raise ValueError("state dict entry %s.* is neither a convolution / linear layer nor a BatchNorm1d" % st)
    if pre_bn is not None:
        raise ValueError("BatchNorm %s has no layer to fold into" % pre_bn[2])
    return [(w, b) for w, b, _, _ in layers]


def _emit_folded(out, layers, ours) -> bool:
    """Folded (w', b') layers -> weight_v = w', weight_g = |w'| (folding it back gives w' exactly), bias, under our names."""
    if len(ours) != len(layers) or any(tuple(w.shape) != vs for (w, _), (_, vs, _) in zip(layers, ours)):
        return False
    for (w, b), (prefix, _, _) in zip(layers, ours):
        v = w.float().contiguous()
        out[prefix + ".bias"] = b.float().contiguous()
        out[prefix + ".weight_g"] = torch.norm_except_dim(v, 2, 0)
        out[prefix + ".weight_v"] = v
    return True


def is_plain_state(state_dict) -> bool:
    """True when some layer holds a plain ``weight`` (no ``weight_g`` / ``weight_v``): a model built without weight-norm."""
    return any(k.endswith(".weight") and state_dict[k].dim() >= 2 and k[:-len(".weight")] + ".weight_v" not in state_dict
               for k in state_dict)


def batchnorm_state_to_weight_norm(state_dict, eps: float = BN_EPS, softplus_nets=()) -> Dict[str, torch.Tensor]:
    """State dict of a model built WITHOUT weight-norm -- plain Conv1d / Linear followed (or, in the pooled head, preceded)
    by ``BatchNorm1d`` (``norm_type="BatchNorm1d"``, the default of python/NNTools.py:27-45,72-115,118-294,517-566 when an
    architecture module has ``weight_norm = False``) -- folded for inference (``_fold_batchnorm_net``) and renamed to this
    package's layer tables.  Eval-mode batch-norm is the affine map ``y = (x - mean) / sqrt(var + eps) * gamma + beta``.  The
    layers come in the same order as in the weight-norm model, so the k-th folded layer of a sub-network is the k-th entry
    of ``conv_keys``.  The activation is not part of a state dict: BatchNorm models are taken to use ReLU, as every BN
    configuration of the reference does (the LayerNorm / Softplus experiment is refused by read_wrapper, which sees the
    modules)."""
    nets = {}
    for k in state_dict:
        nets.setdefault(k.split(".", 1)[0], []).append(k)
    folded_nets = {net: _fold_batchnorm_net(keys, state_dict, eps) for net, keys in nets.items()}
    for cfg in arch.CONFIGS.values():
        if set(cfg.networks()) != set(folded_nets) or set(cfg.softplus_nets) != set(softplus_nets):
            continue
        out = {}
        if all(_emit_folded(out, folded_nets[net], conv_keys(cfg, net)) for net in cfg.networks()):
            return out
    raise ValueError("state dict of a model built without weight-norm does not match any supported HELLO MoE configuration "
                     "(sub-networks with Softplus: %s)" % (sorted(softplus_nets) or "none"))


def init_batchnorm_state(keys_and_shapes, seed: int = 13) -> Dict[str, torch.Tensor]:
    """Deterministic, machine-independent values for a BatchNorm-built reference model, filled in state-dict order: conv /
    linear weights like init_params (uniform, bound 1/sqrt(fan_in)), gamma in [0.5, 1.5], beta in [-0.2, 0.2], running
    means in [-0.5, 0.5] x a layer-dependent scale, running variances in [0.5, 2].  Used by oracle/gen_golden.py (which
    loads it into the reference model) and by the tests (which fold it)."""
    rng = np.random.Generator(np.random.PCG64(seed))
    out = {}
    for key, shape in keys_and_shapes:
        shape = tuple(int(x) for x in shape)
        leaf = key.rsplit(".", 1)[1]
        if leaf == "num_batches_tracked":
            out[key] = torch.tensor(100, dtype=torch.int64)
        elif leaf == "running_var":
            out[key] = torch.from_numpy((0.5 + 1.5 * rng.random(shape)).astype(np.float32))
        elif leaf == "running_mean":
            out[key] = torch.from_numpy(((rng.random(shape) - 0.5) * 4.0).astype(np.float32))
        elif leaf == "weight" and len(shape) == 1:
            out[key] = torch.from_numpy((0.5 + rng.random(shape)).astype(np.float32))
        elif leaf == "bias" and (key.rsplit(".", 1)[0] + ".running_mean") in dict(keys_and_shapes):
            out[key] = torch.from_numpy(((rng.random(shape) - 0.5) * 0.4).astype(np.float32))
        elif leaf == "weight":
            bound = 1.0 / np.sqrt(float(np.prod(shape[1:])))
            out[key] = torch.from_numpy(((rng.random(shape) * 2.0 - 1.0) * bound).astype(np.float32))
        else:                                           # conv / linear bias
            out[key] = torch.from_numpy(((rng.random(shape) * 2.0 - 1.0) * 0.1).astype(np.float32))
    return out


def supported_state(state_dict, softplus_nets=()) -> Dict[str, torch.Tensor]:
    """Any state dict this package can run -> its weight-norm form under MoEAttention names: the live wiring with weight-norm
    (as it is), the legacy wirings (renamed), a model built without weight-norm -- BatchNorm1d layers folded, or no
    normalisation layer at all (``norm_type = "Noop"``, the Softplus configuration).  Anything else raises.  The activation
    is not part of a state dict: the caller names the kinds of sub-network that use Softplus (load_wrapper reads them off
    the modules)."""
    sd = legacy_state_to_attention(state_dict)
    if is_batchnorm_state(sd) or is_plain_state(sd):
        return batchnorm_state_to_weight_norm(sd, softplus_nets=softplus_nets)
    if softplus_nets:
        raise ValueError("no weight-norm configuration with Softplus sub-networks")
    return weight_norm_state(sd)


def init_legacy_state(keys_and_shapes, cfg: arch.ModelConfig, seed: int = 13) -> Dict[str, torch.Tensor]:
    """Deterministic values for a legacy-wiring reference model (``MoEMergedAdvanced``), in its own state-dict order: the
    weight-normed sub-networks get ``init_params(cfg, seed)`` tensor by tensor (same registration order as the live
    sub-network), a sub-network built with BatchNorm1d (the legacy factory's ``meta``) gets ``init_batchnorm_state``."""
    rename = LEGACY_RENAME
    params = init_params(cfg, seed)
    shapes = param_shapes(cfg)
    groups = {}
    for k, shp in keys_and_shapes:
        groups.setdefault(k.split(".", 1)[0], []).append((k, shp))
    out = {}
    for old, ks in groups.items():
        if any(k.endswith(".running_mean") for k, _ in ks):
            out.update(init_batchnorm_state(ks, seed))
        else:
            ours = [k for k in shapes if k.startswith(rename[old] + ".")]
            assert len(ours) == len(ks), (old, len(ours), len(ks))
            for (k, shp), k2 in zip(ks, ours):
                assert tuple(shp) == shapes[k2], (k, k2)
                out[k] = params[k2]
    return {k: out[k] for k, _ in keys_and_shapes}


def cfg_from_state_dict(params: Dict[str, torch.Tensor], softplus_nets=()) -> arch.ModelConfig:
    """Recognise which reference config a MoEAttention state dict belongs to (the activations are the caller's knowledge)."""
    for cfg in arch.CONFIGS.values():
        if set(cfg.softplus_nets) != set(softplus_nets):
            continue
        shapes = param_shapes(cfg)
        if set(shapes) == {k for k in params if not k.endswith(".weight")} or set(shapes) == set(params):
            if all(tuple(params[k].shape) == s for k, s in shapes.items()):
                return cfg
    raise ValueError("state dict does not match any supported HELLO MoE configuration")


class _BlobWriter:
    def __init__(self):
        self.floats: List[np.ndarray] = []
        self.n = 0

    def add(self, t: torch.Tensor) -> int:
        a = t.detach().cpu().contiguous().numpy().astype(np.float32).ravel()
        off = self.n
        pad = (-a.size) % 4           # keep every tensor 16-byte aligned for float4 loads
        self.floats.append(a)
        if pad:
            self.floats.append(np.zeros(pad, np.float32))
        self.n += a.size + pad
        return off


def _conv_rec(conv: arch.Conv, params, prefix, bw: _BlobWriter, act: int = 1) -> List[int]:
    w, b = folded(params, prefix)                      # [cout, cin, k]
    wt = w.permute(2, 1, 0).reshape(conv.k * conv.cin, conv.cout)   # row kk = tap*cin + ci, cout contiguous
    return [conv.cin, conv.cout, conv.k, conv.stride, conv.pad, act if conv.relu else 0, bw.add(wt), bw.add(b)]


def pack_blob(cfg: arch.ModelConfig, params: Dict[str, torch.Tensor]) -> bytes:
    """Serialise folded weights + layer tables for hello_moe_create (include/hello_moe.h).

    Layout (little endian): 128-byte header | records (32 x int32 each) | fp32 data.
      header: magic[8] u32 version u32 n_net_slots u64 rec_off u64 n_rec u64 data_off u64 n_floats
              u32 first_rec[10] u32 n_rec[10]
      record: kind, has_shortcut, conv_a[8], conv_b[8], conv_s[8], pad   with
              conv = cin, cout, k, stride, pad, activation, w_off, b_off   (activation: 0 none, 1 ReLU, 2 Softplus --
                     arch.ACTIVATION_CODES; offsets in floats into the data section)
      conv weights are stored transposed [k*cin, cout] (row = tap*cin + ci) for channel-last activations;
      linear weights as [cout, cin].
    """
    check_params(cfg, params)
    bw = _BlobWriter()
    recs: List[List[int]] = []
    first = [0] * N_NET_SLOTS
    count = [0] * N_NET_SLOTS
    zero8 = [0] * 8
    for net in cfg.networks():
        nid = NET_IDS[net]
        act = arch.ACTIVATION_CODES[cfg.activation(net)]
        first[nid] = len(recs)
        for base, layer in cfg.keyed(net):
            if isinstance(layer, arch.Front):
                continue
            if isinstance(layer, arch.Conv):
                recs.append([KIND_CONV, 0] + _conv_rec(layer, params, base + ".conv1d", bw, act) + zero8 + zero8)
            elif isinstance(layer, arch.MaxPool):
                recs.append([KIND_MAXPOOL, 0] + [0, 0, layer.k, layer.stride, 0, 0, 0, 0] + zero8 + zero8)
            elif isinstance(layer, arch.Res):
                a = _conv_rec(layer.conv_a, params, base + ".ffNetwork.network.0.conv1d", bw, act)
                b = _conv_rec(layer.conv_b, params, base + ".ffNetwork.network.3.conv1d", bw, act)
                s = _conv_rec(layer.conv_s, params, base + ".shNetwork.network.0.conv1d", bw, act) \
                    if layer.conv_shortcut else zero8
                recs.append([KIND_RES, int(layer.conv_shortcut)] + a + b + s)
            elif isinstance(layer, arch.GapLinear):
                w, b = folded(params, arch.linear_key(base))
                recs.append([KIND_GAP_LINEAR, 0] + [layer.cin, layer.cout, 1, 1, 0, 0, bw.add(w), bw.add(b)]
                            + zero8 + zero8)
        count[nid] = len(recs) - first[nid]
    rec_arr = np.zeros((len(recs), _REC_INTS), np.int32)
    for i, r in enumerate(recs):
        rec_arr[i, :len(r)] = r
    data = np.concatenate(bw.floats) if bw.floats else np.zeros(0, np.float32)
    rec_off = _HEADER_BYTES
    data_off = rec_off + rec_arr.nbytes
    header = struct.pack("<8sIIQQQQ", BLOB_MAGIC, BLOB_VERSION, N_NET_SLOTS, rec_off, len(recs), data_off, data.size)
    header += struct.pack("<%dI" % N_NET_SLOTS, *first) + struct.pack("<%dI" % N_NET_SLOTS, *count)
    assert len(header) == _HEADER_BYTES, len(header)
    return header + rec_arr.tobytes() + data.tobytes()
