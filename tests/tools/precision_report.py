#!/usr/bin/env python
"""Error of the three precision modes against the CPU oracle (run on a GPU box): max / mean |dlogit|, max |dP| and the
number of genotype calls that differ at sites whose reference top-2 margin exceeds 2e-3.

    python tests/tools/precision_report.py [n_sites]
"""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from hello_b200 import arch, model, synth, weights          # noqa: E402
from oracle import hello_oracle as O                         # noqa: E402

if __name__ == "__main__":
    n_sites = int(sys.argv[1]) if len(sys.argv) > 1 else 400
    out = {}
    for name in ("single_tech", "hybrid_no_ensemble"):
        cfg = arch.CONFIGS[name]
        params = weights.init_params(cfg, seed=13)
        pl = synth.make_pileups(n_sites, coverage=30, channels=cfg.read_cin, seed=5)
        ref = O.OracleModel(cfg, params).forward(*pl.forward_args())
        post = O.batched_posteriors(cfg, ref, pl.num_alleles_per_site())
        mixed = torch.cat([p[0] for p in post])
        best = torch.tensor([p[3] for p in post], dtype=torch.int32)
        margins = []
        for p in post:
            top = p[0].sort(descending=True).values
            margins.append(float(top[0] - top[1]) if top.numel() > 1 else 1.0)
        clear = torch.tensor(margins) > 2e-3
        for prec in ("fp32", "bf16x3", "bf16"):
            net = model.MoEAttentionB200(cfg, params, device="cuda:0", precision=prec)
            r = net.engine.run(model.DeviceBatch.from_pileups(pl, "cuda:0"))
            head = 0 if cfg.xattn_present[0] else 2
            d = (r.logits[head].cpu() - ref.reshape(-1)).abs()
            differs = (r.best_pair.cpu() != best).any(dim=1)
            out["%s/%s" % (name, prec)] = {"max_abs_dlogit": float(d.max()), "mean_abs_dlogit": float(d.mean()),
                                           "max_abs_dP": float((r.pair_prob[0].cpu() - mixed).abs().max()),
                                           "calls_differ_at_clear_sites": int((differs & clear).sum()),
                                           "calls_differ_total": int(differs.sum()), "sites": n_sites,
                                           "logit_range": [float(ref.min()), float(ref.max())]}
    print(json.dumps(out, indent=1))
