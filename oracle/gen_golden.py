"""Generate tests/golden/*.npz by running the UNMODIFIED reference (imported from /root/reference/python).

Run in the build container only (the reference does not travel to the GPU box):

    PYTHONDONTWRITEBYTECODE=1 python oracle/gen_golden.py

For every supported configuration it
  1. builds the reference model with ``create_moe_attention_model`` from the reference's own config module and
     asserts that hello_b200/arch.py describes it exactly (same state_dict keys, order and shapes);
  2. loads the deterministic weights of hello_b200.weights.init_params into it;
  3. runs the reference's batched ``MoEAttention.forward`` and the per-site
     ``MoEMergedWrapperAdvanced.forward`` (providePredictions=True) on small synthetic pileups;
  4. stores inputs, outputs, the genotype call made with caller_calling.py's rule and a digest of the weights.
One configuration per subprocess: the reference's architecture modules are mutable singletons.
"""
from __future__ import annotations

import importlib
import os
import subprocess
import sys
import warnings

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference/python"
OUT = os.path.join(ROOT, "tests", "golden")

CASES = {
    # name: (n_sites, coverage, seed, uniform_bytes)
    "single_tech": (8, 10, 101, False),
    "single_tech_hp": (6, 8, 102, False),
    "hybrid_no_ensemble": (6, 8, 103, False),
    "hybrid_ensemble2": (6, 8, 104, False),
    "hybrid_full": (6, 8, 105, False),
    "hybrid_no_ensemble_wide": (3, 6, 106, False),
    "single_tech_uniform": (4, 8, 107, True),
    "single_tech_addendum": (5, 8, 108, False),
    "hybrid_no_ensemble_addendum": (4, 6, 109, False),
    # the legacy wiring MoEMergedAdvanced (python/MixtureOfExpertsAdvanced.py:255-484) in its single-technology form, built by
    # createMoEFullMergedAdvancedModel (:614-654) from the legacy architecture modules, useAdditive=True
    "legacy_single_tech": (6, 9, 110, False),
    # the same legacy wiring with two technologies (three experts + meta, additive hybrid features, no ConvCombiners); the
    # factory builds `meta` without weight-norm, i.e. with BatchNorm1d (make_network(configDict, "meta"), :622)
    "legacy_hybrid_additive": (6, 8, 112, False),
    # ... and with both ConvCombiners (alleleConvCombiner / siteConvCombiner = ConvCombinerResNetDeeper, built with BatchNorm1d:
    # the module has no weight-norm switch): the three-expert wiring of MoEAttention, i.e. hello_b200's `hybrid_full`
    "legacy_hybrid_combiners": (6, 8, 113, False),
    # built WITHOUT weight-norm: plain Conv1d / Linear + BatchNorm1d (the architecture modules' default, weight_norm = False),
    # eval mode, deterministic non-trivial running statistics; hello_b200 folds the batch-norms at load
    "single_tech_batchnorm": (6, 9, 111, False),
    # moe_attention_config_single_tech_old_equivalent_layer_norm.py as shipped: norm_type = "Noop", activation = "Softplus"
    # (plain Conv1d / Linear, no normalisation layer, torch.nn.Softplus() after every convolution)
    "single_tech_softplus": (6, 9, 114, False),
}
LEGACY_CONFIG = {"readConvNGS": "MoEReadConvolverDeeper", "alleleConvSingleNGS": "ExpertAlleleConvolverDeeper",
                 "graphConvSingleNGS": "ExpertGraphConvolverDeeper", "weight_norm": True, "kwargs": {"useAdditive": True}}
LEGACY_HYBRID_CONFIG = dict(LEGACY_CONFIG, readConvTGS="MoEReadConvolverDeeper", alleleConvSingleTGS="ExpertAlleleConvolverDeeper",
                            graphConvSingleTGS="ExpertGraphConvolverDeeper", graphConvHybrid="ExpertGraphConvolverDeeper",
                            meta="MetaCombinerDeeper")
LEGACY_COMBINERS_CONFIG = dict(LEGACY_HYBRID_CONFIG, alleleConvCombiner="ConvCombinerResNetDeeper",
                               siteConvCombiner="ConvCombinerResNetDeeper")
LEGACY_CASES = {"legacy_single_tech": ("single_tech", LEGACY_CONFIG), "legacy_hybrid_additive": ("legacy_hybrid_additive", LEGACY_HYBRID_CONFIG),
                "legacy_hybrid_combiners": ("hybrid_full", LEGACY_COMBINERS_CONFIG)}


def run_case(case: str) -> None:
    warnings.filterwarnings("ignore")
    sys.path.insert(0, ROOT)
    sys.path.insert(0, REF)
    import torch
    torch.set_num_threads(1)
    import MixtureOfExpertsAdvanced as M          # the reference
    from hello_b200 import arch, weights, synth

    name = case.replace("_uniform", "").replace("_batchnorm", "")
    name = LEGACY_CASES[name][0] if name in LEGACY_CASES else name
    n_sites, cov, seed, uniform = CASES[case]
    cfg = arch.CONFIGS[name]
    legacy = case.startswith("legacy_")
    batchnorm = case.endswith("_batchnorm")
    name = name.replace("_batchnorm", "")
    cfg = arch.CONFIGS[name]
    plain = bool(cfg.softplus_nets)                       # built without weight-norm and without normalisation layers
    if plain:
        moe = M.create_moe_attention_model(importlib.import_module(arch.REFERENCE_CONFIG_MODULE[name]).configDict).eval()
        for sub_name, sub in moe.named_children():        # arch.py names the sub-networks that got Softplus
            acts = {type(m).__name__ for m in sub.modules()} & {"ReLU", "Softplus"}
            assert acts == ({"Softplus"} if cfg.activation(sub_name) == "softplus" else {"ReLU"}), (sub_name, acts)
    elif batchnorm:
        import architectures.read_convolver as rc_, architectures.compressor_conv_small as cc_, architectures.xattn_subtract as xa_
        for m in (rc_, cc_, xa_):
            m.weight_norm = False
            m.gen_config()
        moe = M.create_moe_attention_model({"read_conv0": rc_.config, "compressor0": cc_.config, "xattn0": xa_.config}).eval()
    elif legacy:
        moe = M.createMoEFullMergedAdvancedModel(dict(LEGACY_CASES[case][1])).eval()
    elif name in arch.REFERENCE_ADDENDUM_MODULE:
        # transfer-learning model: the reference's build_on_top stacks the addendum networks on a trained base model
        # (MixtureOfExpertsDNNFastXferLearning.py:494-502 does this on a DataParallel(WrapperForDataParallel(moe)))
        import types
        import MixtureOfExpertsAdvancedXferLearning as X
        base_name, add_module = arch.REFERENCE_ADDENDUM_MODULE[name]
        base = M.create_moe_attention_model(importlib.import_module(arch.REFERENCE_CONFIG_MODULE[base_name]).configDict)
        add = importlib.import_module(add_module).configDict
        holder = types.SimpleNamespace(module=types.SimpleNamespace(dnn=base))
        moe, _ = X.build_on_top(holder, **{k: X.make_network(add, k) for k in add})
        moe = moe.eval()
    else:
        mod = importlib.import_module(arch.REFERENCE_CONFIG_MODULE[name])
        moe = M.create_moe_attention_model(mod.configDict).eval()
    shapes = weights.param_shapes(cfg)
    sd = moe.state_dict()
    params = weights.init_params(cfg, seed=13)
    bn_keys = None
    if batchnorm or plain:
        bn_keys = [(k, tuple(v.shape)) for k, v in sd.items()]
        bn_state = weights.init_batchnorm_state(bn_keys, seed=13)
        moe.load_state_dict(bn_state)
        moe = moe.eval()                                  # running statistics, not batch statistics
        params = weights.supported_state(moe.state_dict(), softplus_nets=cfg.softplus_nets)
        assert list(params.keys()) == list(shapes.keys())
        assert weights.cfg_from_state_dict(params, softplus_nets=cfg.softplus_nets).name == cfg.name
    elif legacy:
        # the parameters of the live sub-networks in the same registration order under the legacy names (a BatchNorm-built
        # `meta` gets deterministic batch-norm state): hello_b200's mapping must invert this
        legacy_sd = weights.init_legacy_state([(k, tuple(v.shape)) for k, v in sd.items()], cfg, seed=13)
        moe.load_state_dict(legacy_sd)
        moe = moe.eval()
        params = weights.supported_state(moe.state_dict())
        assert list(params.keys()) == list(shapes.keys()), [(a, b) for a, b in zip(params, shapes) if a != b][:4]
        bn_keys = [(k, tuple(v.shape)) for k, v in sd.items()]
    else:
        assert list(sd.keys()) == list(shapes.keys()), "arch.py does not describe the reference model"
        for k in sd:
            assert tuple(sd[k].shape) == shapes[k], k
        moe.load_state_dict(params)
    f_read, f_allele, f_site = arch.flops_model(cfg)

    pl = synth.make_pileups(n_sites, coverage=cov, channels=cfg.read_cin, seed=seed, uniform_bytes=uniform)
    args = pl.forward_args()
    with torch.no_grad():
        res = moe(*args[:3]) if legacy else moe(*args)      # MoEMergedAdvanced.forward(tensors, numAllelesPerSite, numReadsPerAllele)
    out = {
        "digest": np.array(weights.params_digest(params)),
        "torch_version": np.array(torch.__version__),
        "site_allele_off": pl.site_allele_off.numpy(),
        "ref_onehot_idx": pl.ref_onehot.argmax(-1).to(torch.uint8).numpy(),
        "flops": np.array(list(f_read) + [f_allele, f_site], dtype=np.int64),
    }
    if bn_keys is not None:                               # the reference model's own state-dict keys and shapes, in order
        out["bn_keys"] = np.array([k for k, _ in bn_keys])
        out["bn_shapes"] = np.array([",".join(str(x) for x in shp) for _, shp in bn_keys])
    for t, r in enumerate(pl.reads):
        out["reads%d" % t] = r.numpy()
        out["allele_read_off%d" % t] = pl.allele_read_off[t].numpy()
    if cfg.returns_meta:
        experts, meta = res
        out["logits"] = torch.stack([e.reshape(-1) for e in experts]).numpy()
        out["meta"] = meta.numpy()
    else:
        out["logits"] = res.reshape(1, -1).numpy()

    # per-site strict drop-in call, exactly as python/caller_calling.py:651-652, 702-705
    net = M.createMoEFullMergedAdvancedModelWrapper(moe).eval()
    net.providePredictions = True
    mixed, experts, metas, best = [], [[], [], []], [], []
    for s in range(pl.n_sites):
        fd, seg = pl.site_feature_dict(s)
        with torch.no_grad():
            r = net(fd, seg)
        keys = list(r[0].keys())
        mixed += [float(r[0][k]) for k in keys]
        for e in range(3):
            experts[e] += [float(r[1 + e][k]) for k in keys]
        metas.append(r[4].numpy())
        top = sorted([(v, k) for k, v in r[0].items()], reverse=True)[0]
        names = list(fd.keys())
        best.append([names.index(top[1][0]), names.index(top[1][1])])
    out["pair_mixed"] = np.array(mixed, np.float32)
    out["pair_experts"] = np.array(experts, np.float32)
    out["site_meta"] = np.stack(metas).astype(np.float32)
    out["best_pair"] = np.array(best, np.int32)
    os.makedirs(OUT, exist_ok=True)
    np.savez_compressed(os.path.join(OUT, case + ".npz"), **out)
    print("%-28s sites %d alleles %d reads %s logits[%.3f, %.3f] -> %s.npz" % (
        case, pl.n_sites, pl.n_alleles, [int(r.shape[0]) for r in pl.reads],
        out["logits"].min(), out["logits"].max(), case))


if __name__ == "__main__":
    if len(sys.argv) > 1:
        run_case(sys.argv[1])
    else:
        env = dict(os.environ, PYTHONDONTWRITEBYTECODE="1")
        for c in CASES:
            subprocess.run([sys.executable, os.path.abspath(__file__), c], check=True, env=env)
