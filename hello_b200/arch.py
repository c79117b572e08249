"""Layer tables for the sub-networks of HELLO's MoE variant-calling DNN.

The reference describes each sub-network as a Python list of ``{"type", "kwargs"}``
dicts that ``NNTools.Network`` turns into a ``torch.nn.Sequential``
(reference: python/NNTools.py:633-657, python/architectures/*.py).  Here the same
networks are written as short tables of four primitive layer kinds; nothing is
generated from the reference at run time (it does not exist on the GPU box).
``oracle/gen_golden.py`` walks the real reference modules and asserts that these
tables describe them exactly.

Layer kinds
-----------
``Conv``       weight-normed Conv1d + bias (+ ReLU)      -- NNTools.SingleConvLayer, NNTools.py:72-115
``MaxPool``    MaxPool1d(k, stride, pad=0)                -- read_convolver.py:49-56
``Res``        ``relu(conv_b(relu(conv_a(x)))) + sh(x)``  -- NNTools.py:118-294, 569-583 (no ReLU after the add)
``GapLinear``  mean over length, then weight-normed Linear -- NNTools.terminus, NNTools.py:517-566
``Front``      argument plumbing the reference does with Fork/SelectArgument/LinearCombination/
               ConcatenateChannels/Transposer modules; parameter-free, handled by the forward wiring.

``slot`` is the index of the layer inside the reference's ``Sequential`` so that the
reference's ``state_dict`` keys (``<net>.network.<slot>....``) can be derived.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import List, Optional, Tuple, Union

FEATURE_LENGTH = 150  # python/call.py:187; window the C++ encoder emits


@dataclass(frozen=True)
class Conv:
    cin: int
    cout: int
    k: int
    stride: int
    pad: int
    relu: bool = True

    def out_len(self, lin: int) -> int:
        return (lin + 2 * self.pad - self.k) // self.stride + 1

    @property
    def macs_per_pos(self) -> int:
        return self.cin * self.cout * self.k


@dataclass(frozen=True)
class MaxPool:
    k: int
    stride: int

    def out_len(self, lin: int) -> int:
        return (lin - self.k) // self.stride + 1


@dataclass(frozen=True)
class Res:
    """Residual block: conv_a (k3, stride s, pad 1) -> ReLU -> conv_b (k3, s1, p1) -> ReLU, plus shortcut."""
    cin: int
    cout: int
    stride: int
    conv_shortcut: bool  # True: Conv1d(k=1, stride=s, pad=0) with bias; False: identity

    @property
    def conv_a(self) -> Conv:
        return Conv(self.cin, self.cout, 3, self.stride, 1, True)

    @property
    def conv_b(self) -> Conv:
        return Conv(self.cout, self.cout, 3, 1, 1, True)

    @property
    def conv_s(self) -> Optional[Conv]:
        return Conv(self.cin, self.cout, 1, self.stride, 0, False) if self.conv_shortcut else None

    def out_len(self, lin: int) -> int:
        return self.conv_a.out_len(lin)


@dataclass(frozen=True)
class GapLinear:
    cin: int
    cout: int


@dataclass(frozen=True)
class Front:
    n_slots: int


Layer = Union[Conv, MaxPool, Res, GapLinear, Front]

_SLOTS = {Conv: 2, MaxPool: 1, Res: 1, GapLinear: 4}


def with_slots(layers: List[Layer]) -> List[Tuple[int, Layer]]:
    """Pair every layer with its index in the reference's Sequential."""
    out, slot = [], 0
    for layer in layers:
        out.append((slot, layer))
        slot += layer.n_slots if isinstance(layer, Front) else _SLOTS[type(layer)]
    return out


def keyed_layers(net: str, layers: List[Layer], split: Optional[int] = None) -> List[Tuple[str, Layer]]:
    """[(state_dict key prefix of the layer's module, layer)].  Without an addendum the prefix is
    ``<net>.network.<slot>``.  With one (``split`` = index of its first layer) the sub-network is the
    ``torch.nn.Sequential(original, addendum)`` that build_on_top makes
    (MixtureOfExpertsAdvancedXferLearning.py:131-160): ``<net>.0.network.<slot>`` for the original layers and
    ``<net>.1.network.<slot>`` -- slots counted from zero again -- for the added ones."""
    if split is None:
        return [("%s.network.%d" % (net, slot), layer) for slot, layer in with_slots(layers)]
    return ([("%s.0.network.%d" % (net, slot), layer) for slot, layer in with_slots(layers[:split])] +
            [("%s.1.network.%d" % (net, slot), layer) for slot, layer in with_slots(layers[split:])])


def linear_key(prefix: str) -> str:
    """Key prefix of the Linear inside a GapLinear whose first module sits at `prefix` (AdaptiveAvgPool1d, Flatten,
    an empty slot, then the linear: NNTools.terminus, NNTools.py:517-566)."""
    head, slot = prefix.rsplit(".", 1)
    return "%s.%d.linear" % (head, int(slot) + 3)


def read_convolver(cin: int = 6, width: int = 1) -> List[Layer]:
    """architectures/read_convolver.py:9-144 (cin=7: read_convolver_with_hp_channel.py; width=2: _wide)."""
    a, b, c = 16 * width, 32 * width, 64 * width
    return [
        Conv(cin, a, 3, 1, 0), Conv(a, a, 3, 1, 0), Conv(a, b, 3, 1, 0),
        MaxPool(3, 2),
        Res(b, b, 1, False), Res(b, b, 1, False), Res(b, b, 1, False),
        Res(b, c, 2, True),
        Res(c, c, 1, False), Res(c, c, 1, False), Res(c, c, 1, False),
    ]


def compressor(width: int = 1) -> List[Layer]:
    """architectures/compressor_conv_small.py:8-55."""
    a, b = 64 * width, 128 * width
    return [Conv(a, a, 1, 1, 0), Res(a, b, 2, True), Res(b, b, 1, False), Res(b, b, 1, False)]


def _resnet_head(cin: int, n_out: int) -> List[Layer]:
    c = 2 * cin
    return [Conv(cin, cin, 1, 1, 0), Res(cin, c, 2, True), Res(c, c, 1, False), Res(c, c, 1, False),
            GapLinear(c, n_out)]


def xattn(width: int = 1) -> List[Layer]:
    """architectures/xattn_subtract.py:9-95.  Front = Fork[Noop, SelectArgument(1)] + LinearCombination[2,-1]:
    the network body sees ``2*allele - site`` (MixtureOfExpertsAdvanced.py:150-155)."""
    return [Front(2)] + _resnet_head(128 * width, 1)


def meta_convolver() -> List[Layer]:
    """architectures/meta_convolver.py:10-77.  Front = SelectArgument(0): body sees the site-level
    combined features [S,128,18]."""
    return [Front(1)] + _resnet_head(128, 3)


def combiner(width: int = 1) -> List[Layer]:
    """architectures/conv_combiner.py:10-42.  Front = ConcatenateChannels."""
    c = 128 * width
    return [Front(1), Conv(2 * c, 4 * c, 3, 1, 1), Conv(4 * c, c, 1, 1, 0)]


def meta_convolver_ref() -> List[Layer]:
    """architectures/meta_convolver_ref.py:13-106.  Front = SelectArgument(1) + Transposer(1,2): body sees the
    one-hot reference segment as [S,5,150]."""
    return [Front(2), Conv(5, 16, 1, 1, 0),
            Res(16, 32, 2, True), Res(32, 64, 2, True), Res(64, 128, 2, True), Res(128, 256, 2, True),
            GapLinear(256, 3)]


@dataclass(frozen=True)
class ModelConfig:
    """Which sub-networks a model has (reference: python/moe_attention_config_*.py)."""
    name: str
    read_cin: Tuple[int, ...]            # channels of tech0 (and tech1) read tensors
    xattn_present: Tuple[bool, bool, bool]
    combiners: bool
    meta: Optional[str]                  # None | "meta_convolver" | "meta_convolver_ref"
    width: int = 1
    addendum: bool = False               # transfer-learning model: two more residual blocks on top of the read convolvers,
                                         # the compressors and the (single) expert head (architectures/*_addendum.py)
    addendum_blocks: int = 2             # the shipped addenda have two; other depths run too (fused kernels cover <= 2)
    legacy_sum: bool = False             # legacy MoEMergedAdvanced hybrid wiring (useAdditive, no ConvCombiners): the hybrid
                                         # allele feature is compressor0 + compressor1, the hybrid site frame its per-site sum
    softplus_nets: Tuple[str, ...] = ()  # kinds of sub-network ("read_convolver", "compressor", "xattn", ...) whose
                                         # convolutions are followed by torch.nn.Softplus() instead of ReLU: the architecture
                                         # modules' `activation` switch (NNTools.py:72-115) as set by
                                         # moe_attention_config_single_tech_old_equivalent_layer_norm.py -- read convolver and
                                         # expert head honour it, the compressor's generator does not

    def activation(self, net: str) -> str:
        """"relu" or "softplus" for sub-network `net` ("read_convolver0", "xattn2", ...)."""
        return "softplus" if net.rstrip("0123456789") in self.softplus_nets else "relu"

    @property
    def hybrid(self) -> bool:
        return len(self.read_cin) == 2

    @property
    def returns_meta(self) -> bool:
        """True when MoEAttention.forward returns ([e0,e1,e2], meta) (MixtureOfExpertsAdvanced.py:241-248)."""
        return self.hybrid and (self.xattn_present[0] or self.xattn_present[1])

    def networks(self):
        """name -> layer table, in the reference's registration order (MoEAttention.__init__, :104-115)."""
        nets = {}
        extra = lambda c: [Res(c, c, 1, False)] * self.addendum_blocks if self.addendum else []
        for t, cin in enumerate(self.read_cin):
            nets["read_convolver%d" % t] = read_convolver(cin, self.width) + extra(64 * self.width)
        for t in range(len(self.read_cin)):
            nets["compressor%d" % t] = compressor(self.width) + extra(128 * self.width)
        for e in range(3):
            if self.xattn_present[e]:
                x = xattn(self.width)
                # build_on_top strips the pooled linear head of the original expert before stacking
                # (undo_terminating_layers, XferLearning.py:69-91); the addendum ends in a head of its own
                nets["xattn%d" % e] = x[:-1] + extra(256 * self.width) + x[-1:] if self.addendum else x
        if self.meta == "meta_convolver":
            nets["meta"] = meta_convolver()
        elif self.meta == "meta_convolver_ref":
            nets["meta"] = meta_convolver_ref()
        if self.combiners:
            nets["combiner0"] = combiner(self.width)
            nets["combiner1"] = combiner(self.width)
        return nets

    def addendum_split(self, net: str) -> Optional[int]:
        """Index of the first added layer in networks()[net], None when the sub-network has no addendum
        (moe_attention_config_*_addendum.py: read convolvers, compressors and the expert head; never combiners / meta)."""
        if not self.addendum or net.startswith(("combiner", "meta")):
            return None
        base = {"read_convolver": read_convolver, "compressor": compressor}
        for stem, fn in base.items():
            if net.startswith(stem):
                return len(fn())
        return len(xattn()) - 1

    def keyed(self, net: str) -> List[Tuple[str, Layer]]:
        return keyed_layers(net, self.networks()[net], self.addendum_split(net))


CONFIGS = {
    # moe_attention_config_single_tech_old_equivalent_weight_norm.py (Illumina / PacBio models)
    "single_tech": ModelConfig("single_tech", (6,), (True, False, False), False, None),
    # ..._weight_norm_with_hp_channel.py (PacBio haplotagged)
    "single_tech_hp": ModelConfig("single_tech_hp", (7,), (True, False, False), False, None),
    # moe_attention_config_full_hybrid_old_equivalent_weight_norm_no_ensemble.py (shipped hybrid model)
    "hybrid_no_ensemble": ModelConfig("hybrid_no_ensemble", (6, 6), (False, False, True), True, None),
    # ..._ensemble2.py (two experts gated by the reference-segment meta network)
    "hybrid_ensemble2": ModelConfig("hybrid_ensemble2", (6, 6), (True, True, False), False, "meta_convolver_ref"),
    # moe_attention_config_full_hybrid_old_equivalent_weight_norm.py (three experts + meta)
    "hybrid_full": ModelConfig("hybrid_full", (6, 6), (True, True, True), True, "meta_convolver"),
    # ..._no_ensemble_wide.py (2x channels everywhere)
    "hybrid_no_ensemble_wide": ModelConfig("hybrid_no_ensemble_wide", (6, 6), (False, False, True), True, None, 2),
    # legacy wiring MoEMergedAdvanced (python/MixtureOfExpertsAdvanced.py:255-484) with two technologies, useAdditive, no
    # ConvCombiners, built by createMoEFullMergedAdvancedModel (:614-654) from MoEReadConvolverDeeper / ExpertAlleleConvolverDeeper
    # / ExpertGraphConvolverDeeper / MetaCombinerDeeper: three experts on 2a - s, meta on the summed site frame
    "legacy_hybrid_additive": ModelConfig("legacy_hybrid_additive", (6, 6), (True, True, True), False, "meta_convolver",
                                          legacy_sum=True),
    # moe_attention_config_single_tech_old_equivalent_weight_norm_addendum.py stacked on the single-tech model by
    # MixtureOfExpertsAdvancedXferLearning.build_on_top (:94-183)
    "single_tech_addendum": ModelConfig("single_tech_addendum", (6,), (True, False, False), False, None, 1, True),
    # ..._full_hybrid_old_equivalent_weight_norm_no_ensemble_addendum.py stacked on the shipped hybrid model
    "hybrid_no_ensemble_addendum": ModelConfig("hybrid_no_ensemble_addendum", (6, 6), (False, False, True), True, None, 1,
                                               True),
    # moe_attention_config_single_tech_old_equivalent_layer_norm.py: the single-technology networks built with
    # `norm_type = "Noop"` and `activation = "Softplus"`: plain Conv1d / Linear without normalisation layers (one
    # BatchNorm1d survives in the expert's pooled head and is folded), Softplus in the read convolver and the expert head
    "single_tech_softplus": ModelConfig("single_tech_softplus", (6,), (True, False, False), False, None,
                                        softplus_nets=("read_convolver", "xattn")),
}

ACTIVATION_CODES = {"relu": 1, "softplus": 2}      # the `relu` field of a convolution record in the weight blob (0 = none)

# configs whose model is <base model> + build_on_top(<addendum config module>)
REFERENCE_ADDENDUM_MODULE = {
    "single_tech_addendum": ("single_tech", "moe_attention_config_single_tech_old_equivalent_weight_norm_addendum"),
    "hybrid_no_ensemble_addendum": ("hybrid_no_ensemble",
                                    "moe_attention_config_full_hybrid_old_equivalent_weight_norm_no_ensemble_addendum"),
}

REFERENCE_CONFIG_MODULE = {
    "single_tech": "moe_attention_config_single_tech_old_equivalent_weight_norm",
    "single_tech_hp": "moe_attention_config_single_tech_old_equivalent_weight_norm_with_hp_channel",
    "hybrid_no_ensemble": "moe_attention_config_full_hybrid_old_equivalent_weight_norm_no_ensemble",
    "hybrid_ensemble2": "moe_attention_config_full_hybrid_old_equivalent_weight_norm_ensemble2",
    "hybrid_full": "moe_attention_config_full_hybrid_old_equivalent_weight_norm",
    "hybrid_no_ensemble_wide": "moe_attention_config_full_hybrid_old_equivalent_weight_norm_no_ensemble_wide",
    "single_tech_softplus": "moe_attention_config_single_tech_old_equivalent_layer_norm",
}


def net_macs(layers: List[Layer], lin: int) -> int:
    """Multiply-accumulates for one item through a layer table (bias/ReLU/pool not counted)."""
    total, length = 0, lin
    for layer in layers:
        if isinstance(layer, Conv):
            length = layer.out_len(length)
            total += layer.macs_per_pos * length
        elif isinstance(layer, MaxPool):
            length = layer.out_len(length)
        elif isinstance(layer, Res):
            lo = layer.out_len(length)
            total += (layer.conv_a.macs_per_pos + layer.conv_b.macs_per_pos) * lo
            if layer.conv_shortcut:
                total += layer.conv_s.macs_per_pos * lo
            length = lo
        elif isinstance(layer, GapLinear):
            total += layer.cin * layer.cout
            length = 1
    return total


def net_out_shape(layers: List[Layer], lin: int) -> Tuple[int, int]:
    """(channels, length) coming out of a layer table."""
    length, ch = lin, None
    for layer in layers:
        if isinstance(layer, Front):
            continue
        if isinstance(layer, GapLinear):
            return layer.cout, 1
        length = layer.out_len(length)
        if not isinstance(layer, MaxPool):
            ch = layer.cout
    return ch, length


def flops_model(cfg: ModelConfig):
    """(F_read per tech, F_allele, F_site) in FLOPs, dead site-compressor branch excluded."""
    nets = cfg.networks()
    f_read = tuple(2 * net_macs(nets["read_convolver%d" % t], FEATURE_LENGTH) for t in range(len(cfg.read_cin)))
    _, l_read = net_out_shape(nets["read_convolver0"], FEATURE_LENGTH)
    _, l_comp = net_out_shape(nets["compressor0"], l_read)
    f_allele = sum(2 * net_macs(nets["compressor%d" % t], l_read) for t in range(len(cfg.read_cin)))
    f_site = 0
    for e in range(3):
        if cfg.xattn_present[e]:
            f_allele += 2 * net_macs(nets["xattn%d" % e], l_comp)
    if cfg.combiners:
        f_allele += 2 * net_macs(nets["combiner0"], l_comp)
        f_site += 2 * net_macs(nets["combiner1"], l_comp)
    if cfg.meta == "meta_convolver":
        f_site += 2 * net_macs(nets["meta"], l_comp)
    elif cfg.meta == "meta_convolver_ref":
        f_site += 2 * net_macs(nets["meta"], FEATURE_LENGTH)
    return f_read, f_allele, f_site
