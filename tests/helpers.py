"""Shared helpers for the parity tests."""
import os

import numpy as np
import torch

from hello_b200 import arch, synth, weights

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
GOLDEN_CASES = ["single_tech", "single_tech_hp", "hybrid_no_ensemble", "hybrid_ensemble2", "hybrid_full",
                "hybrid_no_ensemble_wide", "single_tech_uniform"]


def load_golden(case):
    g = dict(np.load(os.path.join(GOLDEN, case + ".npz")))
    cfg = arch.CONFIGS[case.replace("_uniform", "")]
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
