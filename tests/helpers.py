"""Shared helpers for the parity tests."""
import os

import numpy as np
import torch

from hello_b200 import arch, synth, weights

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
GOLDEN_CASES = ["single_tech", "single_tech_hp", "hybrid_no_ensemble", "hybrid_ensemble2", "hybrid_full",
                "hybrid_no_ensemble_wide", "single_tech_uniform", "single_tech_addendum", "hybrid_no_ensemble_addendum",
                "legacy_single_tech"]


def batchnorm_params(case="single_tech_batchnorm"):
    """The folded parameters of the BatchNorm-built golden model: its reference state dict is re-created from the key /
    shape list stored in the fixture (weights.init_batchnorm_state is deterministic) and folded by
    weights.batchnorm_state_to_weight_norm -- the same two steps a user's BatchNorm model goes through in from_state_dict."""
    g = np.load(os.path.join(GOLDEN, case + ".npz"))
    keys = [(str(k), tuple(int(x) for x in str(s).split(",") if x)) for k, s in zip(g["bn_keys"], g["bn_shapes"])]
    state = weights.init_batchnorm_state(keys, seed=13)
    cfg = arch.CONFIGS.get(case)                       # the Softplus configuration names its Softplus sub-networks
    return state, weights.supported_state(state, softplus_nets=cfg.softplus_nets if cfg else ())


LEGACY_CASE_CFG = {"legacy_single_tech": "single_tech", "legacy_hybrid_combiners": "hybrid_full"}


def legacy_params(case="legacy_hybrid_additive"):
    """(legacy state dict, folded / renamed parameters) of a legacy-wiring golden model: the reference model's state dict is
    re-created from the key / shape list stored in the fixture (weights.init_legacy_state is deterministic) and goes through
    weights.supported_state -- what from_state_dict does with a user's legacy model."""
    g = np.load(os.path.join(GOLDEN, case + ".npz"))
    keys = [(str(k), tuple(int(x) for x in str(s).split(",") if x)) for k, s in zip(g["bn_keys"], g["bn_shapes"])]
    state = weights.init_legacy_state(keys, arch.CONFIGS[LEGACY_CASE_CFG.get(case, case)], seed=13)
    return state, weights.supported_state(state)


def load_golden(case):
    g = dict(np.load(os.path.join(GOLDEN, case + ".npz")))
    cfg = arch.CONFIGS[LEGACY_CASE_CFG.get(case, case).replace("_uniform", "").replace("_batchnorm", "")]
    reads, offs = [], []
    for t in range(len(cfg.read_cin)):
        reads.append(torch.from_numpy(g["reads%d" % t]))
        offs.append(torch.from_numpy(g["allele_read_off%d" % t]))
    onehot = torch.nn.functional.one_hot(torch.from_numpy(g["ref_onehot_idx"]).long(), 5).float()
    pl = synth.Pileups(tuple(reads), tuple(offs), torch.from_numpy(g["site_allele_off"]), onehot)
    return cfg, pl, g


_params_cache = {}


def params_for(cfg, seed=13):
    key = (cfg.name, seed)
    if key not in _params_cache:
        _params_cache[key] = weights.init_params(cfg, seed=seed)
    return _params_cache[key]


def flat_result(cfg, res):
    """Batched forward result -> (logits [n_heads, A], meta or None)."""
    if cfg.returns_meta:
        experts, meta = res
        return torch.stack([e.reshape(-1) for e in experts]), meta
    return res.reshape(1, -1), None


def readconv_phase_reference(cfg, params, reads_rlc, tech=0):
    """Post-activation fp32 values of the 17 layer phases of the read convolver (oracle layers), each [R, C, L]:
    0-1 stem convs, 2 stem conv 3 + max-pool, then (conv_a, block output) of the seven residual blocks."""
    import torch.nn.functional as F
    from oracle import hello_oracle as O
    net = O.OracleModel(cfg, params).nets["read_convolver%d" % tech]
    L = net.layers
    x = reads_rlc.transpose(1, 2).float()
    outs = []
    with torch.no_grad():
        x = net._conv(x, L[0][1], L[0][2]); outs.append(x)
        x = net._conv(x, L[1][1], L[1][2]); outs.append(x)
        x = F.max_pool1d(net._conv(x, L[2][1], L[2][2]), 3, 2, 0); outs.append(x)
        for li in range(4, 11):
            _, layer, (wa, wb, ws) = L[li]
            t = net._conv(x, layer.conv_a, wa); outs.append(t)
            sh = net._conv(x, layer.conv_s, ws) if ws is not None else x
            x = net._conv(t, layer.conv_b, wb) + sh; outs.append(x)
    return outs


def readconv_phase_from_dump(dump, phase, n_reads, length):
    """Undo the kernel's row packing: dump [groups, 512, 64] -> [R, C, L] for one phase (include/hello_moe.h)."""
    out_ch = 16 if phase < 2 else (32 if phase < 9 else 64)
    out = torch.zeros((n_reads, out_ch, length))
    for r in range(n_reads):
        g, i = divmod(r, 3)
        if 2 <= phase < 9:
            # 32-channel stage, space-to-depth: pitch 40, row q holds positions 2q (slots 0-31) and 2q+1 (slots 32-63)
            rows = dump[g, i * 40:i * 40 + (length + 1) // 2, :64]                       # [ceil(L/2), 64]
            both = rows.reshape(-1, 2, 32).reshape(-1, 32)[:length]                       # position-major
            out[r] = both.t()
        else:
            pitch = 160 if phase < 2 else 40
            out[r] = dump[g, i * pitch:i * pitch + length, :out_ch].t()
    return out


def head_phase_reference(cfg, params, net_name, x_lc):
    """Post-activation fp32 values of the 7 layer phases of a fused head network (oracle layers), each [n, C, L]:
    0 the 1x1 conv, then (conv_a, block output) of the three residual blocks.  x_lc: [n, L, C] channel-last."""
    from oracle import hello_oracle as O
    net = O.OracleModel(cfg, params).nets[net_name]
    L = net.layers
    x = x_lc.transpose(1, 2).float()
    outs = []
    with torch.no_grad():
        x = net._conv(x, L[0][1], L[0][2]); outs.append(x)
        for li in range(1, 4):
            _, layer, (wa, wb, ws) = L[li]
            t = net._conv(x, layer.conv_a, wa); outs.append(t)
            sh = net._conv(x, layer.conv_s, ws) if ws is not None else x
            x = net._conv(t, layer.conv_b, wb) + sh; outs.append(x)
    return outs


def head_phase_from_dump(dump, phase, n_items, cin, ch, length):
    """Undo the head kernel's row packing: dump [groups, 256, 256] -> [n, C, L] for one phase (include/hello_moe.h)."""
    per = 6 if cin == 64 else 12
    pitch = (40 if cin == 64 else 20) // (1 if phase == 0 else 2)
    out = torch.zeros((n_items, ch, length))
    for r in range(n_items):
        g, i = divmod(r, per)
        out[r] = dump[g, i * pitch:i * pitch + length, :ch].t()
    return out
