"""The CPU oracle against golden vectors produced by the unmodified reference (oracle/gen_golden.py)."""
import numpy as np
import pytest
import torch

from helpers import GOLDEN_CASES, flat_result, load_golden, params_for
from hello_b200 import arch, weights
from oracle import hello_oracle as O

# the oracle is bit-identical to the reference on the machine that made the fixtures; a different host CPU may
# pick another oneDNN/ATen kernel, so allow fp32 reassociation noise
TOL = 5e-5


@pytest.mark.parametrize("case", GOLDEN_CASES)
def test_weights_are_the_ones_the_reference_was_given(case):
    cfg, _, g = load_golden(case)
    assert weights.params_digest(params_for(cfg)) == str(g["digest"])


@pytest.mark.parametrize("case", GOLDEN_CASES)
def test_flops_model(case):
    cfg, _, g = load_golden(case)
    f_read, f_allele, f_site = arch.flops_model(cfg)
    assert list(f_read) + [f_allele, f_site] == g["flops"].tolist()


@pytest.mark.parametrize("case", GOLDEN_CASES)
def test_batched_forward_matches_reference(case):
    cfg, pl, g = load_golden(case)
    torch.set_num_threads(1)
    res = O.OracleModel(cfg, params_for(cfg)).forward(*pl.forward_args())
    logits, meta = flat_result(cfg, res)
    np.testing.assert_allclose(logits.numpy(), g["logits"], rtol=0, atol=TOL)
    if meta is not None:
        np.testing.assert_allclose(meta.numpy(), g["meta"], rtol=0, atol=TOL)


@pytest.mark.parametrize("case", GOLDEN_CASES)
def test_per_site_wrapper_matches_reference(case):
    cfg, pl, g = load_golden(case)
    torch.set_num_threads(1)
    model = O.OracleModel(cfg, params_for(cfg))
    mixed, experts, metas, best = [], [[], [], []], [], []
    for s in range(pl.n_sites):
        fd, seg = pl.site_feature_dict(s)
        r = O.wrapper_forward(model, fd, seg, provide_predictions=True)
        keys = list(r[0].keys())
        names = list(fd.keys())
        assert keys == [(names[i], names[j]) for i, j in O.pair_list(len(names))]
        mixed += [float(r[0][k]) for k in keys]
        for e in range(3):
            experts[e] += [float(r[1 + e][k]) for k in keys]
        metas.append(r[4].numpy())
        key, value, qual = O.call_genotype(r[0])
        best.append([names.index(key[0]), names.index(key[1])])
        assert 0 <= qual <= 80.0001
    np.testing.assert_allclose(np.array(mixed, np.float32), g["pair_mixed"], rtol=0, atol=TOL)
    np.testing.assert_allclose(np.array(experts, np.float32), g["pair_experts"], rtol=0, atol=TOL)
    np.testing.assert_allclose(np.stack(metas), g["site_meta"], rtol=0, atol=TOL)
    assert np.array_equal(np.array(best, np.int32), g["best_pair"])


def test_batchnorm_model_is_folded_like_the_reference_evaluates_it():
    """tests/golden/single_tech_batchnorm.npz: the reference model built WITHOUT weight-norm (Conv1d -> BatchNorm1d -> ReLU,
    BatchNorm1d -> Linear in the pooled head; python/NNTools.py:72-115,118-294,517-566) in eval mode.  Folding its state dict
    (weights.batchnorm_state_to_weight_norm) must give the parameters the fixture was made with and the oracle must
    reproduce the reference's logits and per-site probabilities."""
    from helpers import batchnorm_params
    cfg, pl, g = load_golden("single_tech_batchnorm")
    state, params = batchnorm_params()
    assert weights.is_batchnorm_state(state) and not weights.is_batchnorm_state(params)
    assert weights.cfg_from_state_dict(params).name == "single_tech"
    assert weights.params_digest(params) == str(g["digest"])
    with pytest.raises(ValueError, match="without weight-norm"):
        weights.weight_norm_state(state)                  # the strict weight-norm reader still refuses it
    torch.set_num_threads(1)
    model = O.OracleModel(cfg, params)
    res = model.forward(*pl.forward_args())
    np.testing.assert_allclose(res.reshape(1, -1).numpy(), g["logits"], rtol=0, atol=TOL)
    mixed = []
    for s in range(pl.n_sites):
        fd, seg = pl.site_feature_dict(s)
        r = O.wrapper_forward(model, fd, seg, provide_predictions=True)
        mixed += [float(v) for v in r[0].values()]
    np.testing.assert_allclose(np.array(mixed, np.float32), g["pair_mixed"], rtol=0, atol=TOL)


def test_softplus_configuration_matches_the_reference():
    """tests/golden/single_tech_softplus.npz: moe_attention_config_single_tech_old_equivalent_layer_norm.py as shipped
    (norm_type "Noop", activation "Softplus": plain Conv1d / Linear, torch.nn.Softplus() in the read convolver and the expert
    head, ReLU in the compressor, one BatchNorm1d left in the pooled head).  The state dict says nothing about activations:
    the caller names the Softplus sub-networks; the oracle then reproduces the reference's logits and probabilities."""
    from helpers import batchnorm_params
    cfg, pl, g = load_golden("single_tech_softplus")
    assert cfg.activation("read_convolver0") == "softplus" and cfg.activation("compressor0") == "relu" and cfg.activation("xattn0") == "softplus"
    state, params = batchnorm_params("single_tech_softplus")
    assert weights.is_plain_state(state) and not weights.is_plain_state(params)
    assert weights.cfg_from_state_dict(params, softplus_nets=cfg.softplus_nets).name == "single_tech_softplus"
    assert weights.cfg_from_state_dict(params).name == "single_tech"        # same tensors: only the caller knows the activation
    assert weights.params_digest(params) == str(g["digest"])
    torch.set_num_threads(1)
    model = O.OracleModel(cfg, params)
    res = model.forward(*pl.forward_args())
    np.testing.assert_allclose(res.reshape(1, -1).numpy(), g["logits"], rtol=0, atol=TOL)
    relu_res = O.OracleModel(arch.CONFIGS["single_tech"], params).forward(*pl.forward_args())
    assert (relu_res.reshape(1, -1).numpy() - g["logits"]).__abs__().max() > 0.05           # the activation matters
    mixed = []
    for s in range(pl.n_sites):
        fd, seg = pl.site_feature_dict(s)
        r = O.wrapper_forward(model, fd, seg, provide_predictions=True)
        mixed += [float(v) for v in r[0].values()]
    np.testing.assert_allclose(np.array(mixed, np.float32), g["pair_mixed"], rtol=0, atol=TOL)


@pytest.mark.parametrize("case", ["legacy_hybrid_additive", "legacy_hybrid_combiners"])
def test_legacy_hybrid_wiring_matches_the_reference(case):
    """tests/golden/legacy_hybrid_additive.npz: MoEMergedAdvanced (python/MixtureOfExpertsAdvanced.py:255-484) with two
    technologies built by createMoEFullMergedAdvancedModel (:614-654): three experts on 2a - s, the hybrid allele feature the
    SUM of the two technologies' (:408-412), its site frame the per-site sum of that (:425-436), meta (BatchNorm-built by the
    factory) on it.  The state dict is renamed / folded by weights.supported_state; the oracle reproduces logits, meta weights
    and the per-site wrapper outputs of the reference.
    tests/golden/legacy_hybrid_combiners.npz: the same factory with alleleConvCombiner / siteConvCombiner =
    ConvCombinerResNetDeeper (BatchNorm-built, folded): the hybrid allele feature and site frame come from the combiners
    (:408-436) -- MoEAttention's three-expert wiring, `hybrid_full`."""
    from helpers import legacy_params
    cfg, pl, g = load_golden(case)
    assert cfg.legacy_sum == (case == "legacy_hybrid_additive") and cfg.combiners != cfg.legacy_sum and cfg.returns_meta
    state, params = legacy_params(case)
    assert any(k.startswith("readConv1.") for k in state) and weights.is_batchnorm_state(state)
    assert any(k.startswith("siteConvCombiner.") for k in state) == cfg.combiners
    assert weights.cfg_from_state_dict(params).name == cfg.name
    assert weights.params_digest(params) == str(g["digest"])
    torch.set_num_threads(1)
    model = O.OracleModel(cfg, params)
    res = model.forward(*pl.forward_args())
    logits, meta = flat_result(cfg, res)
    np.testing.assert_allclose(logits.numpy(), g["logits"], rtol=0, atol=TOL)
    np.testing.assert_allclose(meta.numpy(), g["meta"], rtol=0, atol=TOL)
    mixed, best = [], []
    for s in range(pl.n_sites):
        fd, seg = pl.site_feature_dict(s)
        r = O.wrapper_forward(model, fd, seg, provide_predictions=True)
        mixed += [float(v) for v in r[0].values()]
        key, _, _ = O.call_genotype(r[0])
        names = list(fd.keys())
        best.append([names.index(key[0]), names.index(key[1])])
    np.testing.assert_allclose(np.array(mixed, np.float32), g["pair_mixed"], rtol=0, atol=TOL)
    assert np.array_equal(np.array(best, np.int32), g["best_pair"])


def test_batched_equals_per_site_tail():
    cfg, pl, g = load_golden("hybrid_full")
    res = O.OracleModel(cfg, params_for(cfg)).forward(*pl.forward_args())
    post = O.batched_posteriors(cfg, res, pl.num_alleles_per_site())
    mixed = torch.cat([p[0] for p in post]).numpy()
    np.testing.assert_allclose(mixed, g["pair_mixed"], rtol=0, atol=TOL)


def test_reduce_slots_is_a_segmented_sum():
    d = torch.randn(11, 3, 5)
    slots = [1, 4, 2, 1, 3]
    out = O.reduce_slots(d, slots)
    off = np.cumsum([0] + slots)
    for g_, (a, b) in enumerate(zip(off[:-1], off[1:])):
        torch.testing.assert_close(out[g_], d[a:b].sum(0), rtol=1e-5, atol=1e-5)


def test_call_rule_tie_break_and_quality():
    # equal values: the lexicographically greatest key wins (sorted(..., reverse=True)[0])
    key, value, qual = O.call_genotype({("A", "A"): 0.25, ("A", "C"): 0.5, ("C", "C"): 0.5})
    assert key == ("C", "C") and value == 0.5
    key, value, qual = O.call_genotype({("A", "A"): 1.0})
    assert abs(qual - 80.0) < 1e-6
    assert O.remix_float64([[0.5], [0.25], [0.125]], [0.5, 0.25, 0.25]) == [0.5 * 0.5 + 0.25 * 0.25 + 0.125 * 0.25]


def test_final_calls_match_reference_vcfrecords():
    """oracle.final_calls against records made by the reference's own prepareVcf.vcfRecords
    (oracle/gen_final_calls.py): top allele pair of expert0/1/2, best, mean; QUAL; np.argmax(meta)."""
    import os
    from helpers import GOLDEN
    from oracle import hello_oracle as O
    g = np.load(os.path.join(GOLDEN, "final_calls.npz"))
    names_all = [str(x) for x in g["names"]]
    p0 = 0
    for s, n in enumerate(g["n_alleles"]):
        n = int(n)
        names = [names_all[i] for i in g["allele_name_idx"][s, :n]]
        pairs = [(names[i], names[j]) for i in range(n) for j in range(i, n)]
        preds = [{pairs[q]: g["experts"][e, p0 + q] for q in range(len(pairs))} for e in range(3)]
        got = O.final_calls(preds, g["meta"][s])
        assert got["choice"] == int(g["best_expert"][s])
        for c, key in enumerate(str(x) for x in g["calls"]):
            top, qual = got[key]
            assert (names.index(top[0]), names.index(top[1])) == tuple(g["call_pair"][s, c]), (s, key)
            assert abs(qual - g["call_qual"][s, c]) <= 1e-12 * max(1.0, abs(qual)), (s, key)
        p0 += len(pairs)
