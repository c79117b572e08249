"""CUDA path (through the C ABI) against the CPU oracle and the reference-made golden vectors."""
import numpy as np
import pytest
import torch

from helpers import (GOLDEN_CASES, flat_result, head_phase_from_dump, head_phase_reference, load_golden, params_for,
                     readconv_phase_from_dump, readconv_phase_reference)
from hello_b200 import arch, synth, weights

pytestmark = pytest.mark.gpu

# Floating-point tolerance (north_star: "max abs error <= 1e-3 in fp32 accumulate").  fp32 mode is plain fp32
# FMA with a different summation order than oneDNN's, so it lands far inside that.
# bf16x3 runs the read convolver on tcgen05 with hi+lo bf16 operands (3 MMAs, fp32 accumulate): ~1e-5 relative per
# layer.  bf16 (single MMA) is the "fast" mode, reported separately and only checked loosely.
TOL_LOGIT = {"fp32": 2e-4, "bf16x3": 1e-3}
TOL_PROB = {"fp32": 5e-5, "bf16x3": 5e-4}
# north_star speaks of log-posteriors: pair probabilities above LOG_FLOOR must also agree in log space (an absolute bound
# alone would let p = 1e-6 be off by a factor of 500).  |d log P| <= (number of alleles) * |d logit|, so with the logit
# tolerances above and up to ~6 alleles per site:
TOL_LOGP = {"fp32": 2e-3, "bf16x3": 6e-3}
LOG_FLOOR = 1e-6
# Sites whose reference top-2 margin is below twice the probability tolerance are not required to give the same call
# (CPU and GPU exp/log/sigmoid are not bit-identical); their share must stay small and is printed.
MAX_TIGHT_SHARE = 0.02
PRECISIONS = ["bf16x3", "fp32"]          # the shipped default first
TC_LAYER_REL = {"bf16x3": 5e-5, "bf16": 4e-2}
DEV = "cuda:0"


@pytest.fixture(scope="module")
def gpu():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from hello_b200 import model
    return model


_engines = {}


def net_for(model, cfg, precision="bf16x3", **kw):
    key = (cfg.name, precision, tuple(sorted(kw.items())))
    if key not in _engines:
        _engines[key] = model.MoEAttentionB200(cfg, params_for(cfg), device=DEV, precision=precision, **kw)
    return _engines[key]


def oracle_for(cfg):
    from oracle import hello_oracle as O
    return O.OracleModel(cfg, params_for(cfg))


# ------------------------------------------------------------------------------------------------ sub-networks
@pytest.mark.parametrize("name", ["single_tech", "single_tech_hp"])
def test_read_convolver_per_read(gpu, name):
    cfg = arch.CONFIGS[name]
    pl = synth.make_pileups(5, coverage=9, channels=cfg.read_cin, seed=21)
    net = net_for(gpu, cfg, "fp32")              # the CUDA-core path; the tensor-core path has its own per-layer tests
    from hello_b200 import _lib
    got_rlc = net.engine.run_net("read_convolver0", pl.reads[0], _lib.LAYOUT_RLC).cpu()
    got_rcl = net.engine.run_net("read_convolver0", pl.reads[0].transpose(1, 2).contiguous(), _lib.LAYOUT_RCL).cpu()
    ref = oracle_for(cfg).read_features(pl.reads[0].transpose(1, 2))            # [R, 64, 36]
    assert got_rlc.shape == (pl.reads[0].shape[0], 36, 64)
    # (row-major rows take the first convolution through conv_stem_u8_kernel, channel-major ones through conv1d_fp32_kernel:
    # the two accumulate in the same order and must agree bit for bit)
    assert torch.equal(got_rlc, got_rcl), "the two input layouts must give identical results"
    err = (got_rlc.transpose(1, 2) - ref).abs().max().item()
    assert err < 2e-3 * max(1.0, ref.abs().max().item() / 100), err


@pytest.mark.parametrize("net_name,shape", [("compressor0", (7, 36, 64)), ("xattn0", (7, 18, 128))])
def test_head_networks(gpu, net_name, shape):
    cfg = arch.CONFIGS["single_tech"]
    g = torch.Generator().manual_seed(5)
    x = (torch.randn(shape, generator=g) * 40).float()
    got = net_for(gpu, cfg, "fp32").engine.run_net(net_name, x).cpu()
    ref = oracle_for(cfg).nets[net_name](x.transpose(1, 2))
    if ref.dim() == 3:
        ref = ref.transpose(1, 2)
    else:
        got = got.reshape(ref.shape)
    scale = max(1.0, ref.abs().max().item())
    assert (got - ref).abs().max().item() < 2e-5 * scale


@pytest.mark.parametrize("precision", ["bf16x3", "bf16"])
def test_generic_tensor_core_layers(gpu, precision):
    """Sub-networks without a fused kernel run layer by layer on the generic tcgen05 conv kernel in the tensor-core
    modes: meta_convolver_ref (stride-2 blocks 16 -> 256 channels on the one-hot reference) and the 2x-wide model's
    compressor / xattn / combiner, for item counts that leave partial 128-row tiles."""
    tol = TC_LAYER_REL[precision] * (40 if precision == "bf16" else 4)
    cfg = arch.CONFIGS["hybrid_ensemble2"]
    g = torch.Generator().manual_seed(31)
    eng, orc = net_for(gpu, cfg, precision).engine, oracle_for(cfg)
    for n in (1, 7, 130):
        onehot = torch.nn.functional.one_hot(torch.randint(0, 5, (n, 150), generator=g), 5).float()
        before = eng.launch_count()
        got = eng.run_net("meta", onehot).cpu().reshape(n, 3)
        assert eng.launch_count() - before == 14          # 13 convolutions + pooled linear head
        ref = orc.nets["meta"](onehot.transpose(1, 2))
        assert (got - ref).abs().max().item() < tol * max(1.0, ref.abs().max().item()), n
    cfgw = arch.CONFIGS["hybrid_no_ensemble_wide"]
    engw, orcw = net_for(gpu, cfgw, precision).engine, oracle_for(cfgw)
    for net_name, shape in (("compressor0", (9, 36, 128)), ("xattn2", (9, 18, 256)), ("combiner0", (9, 18, 512))):
        x = (torch.randn(shape, generator=g) * 10).float()
        got = engw.run_net(net_name, x).cpu()
        ref = orcw.nets[net_name](x.transpose(1, 2))
        ref = ref.transpose(1, 2) if ref.dim() == 3 else ref.reshape(got.shape)
        assert (got - ref).abs().max().item() < tol * max(1.0, ref.abs().max().item()), net_name


def test_standard_models_run_on_the_fused_kernels(gpu):
    """Kernel launches per chunk in the tensor-core mode: a silent fall-back to per-layer kernels would show here."""
    expect = {"single_tech": 6,            # read convolver (+ allele sum), compressor, segsum, xattn, site index + posterior
              "hybrid_no_ensemble": 11,    # 2 x (read convolver, compressor, segsum) + 2 combiners + xattn2 + 2
              # addendum models (two more residual blocks per sub-network): the same fused kernels, more phases
              "single_tech_addendum": 6,
              "hybrid_no_ensemble_addendum": 11}
    for name, n_launch in expect.items():
        cfg = arch.CONFIGS[name]
        eng = net_for(gpu, cfg, "bf16x3").engine
        # a batch that fills the GPU: the read convolver sums the alleles itself; a handful of sites (fewer than four work
        # items per SM): per-read maps + segsum_kernel, one more launch per technology (every SM gets an item instead of
        # one SM walking a whole allele)
        for n_sites, cov, extra in ((400, 20, 0), (6, 6, len(cfg.read_cin))):
            pl = synth.make_pileups(n_sites, coverage=cov, channels=cfg.read_cin, seed=2)
            assert (pl.reads[0].shape[0] > 4 * 148 * 9) == (extra == 0)
            batch = gpu.DeviceBatch.from_pileups(pl, DEV)
            before = eng.launch_count()
            eng.run(batch)
            assert eng.launch_count() - before == n_launch + 1 + extra, (name, n_sites)   # + fill_meta_default (no meta network)


def test_small_and_large_batches_agree_bit_for_bit(gpu):
    """The same sites scored alone (small-batch path: per-read maps + segsum) and inside a batch that fills the GPU (allele
    sum fused into the read convolver) give bit-identical logits: both add an allele's reads in read order."""
    cfg = arch.CONFIGS["single_tech"]
    net = net_for(gpu, cfg, "bf16x3")
    pl = synth.make_pileups(400, coverage=20, channels=cfg.read_cin, seed=8)
    whole = net.forward(*pl.forward_args()).cpu().reshape(-1)
    sao = pl.site_allele_off
    a1 = int(sao[5])
    r1 = int(pl.allele_read_off[0][a1])
    sub = synth.Pileups((pl.reads[0][:r1],), (pl.allele_read_off[0][:a1 + 1],), sao[:6], pl.ref_onehot[:5])
    part = net.forward(*sub.forward_args()).cpu().reshape(-1)
    assert torch.equal(part, whole[:part.numel()])


@pytest.mark.parametrize("precision", PRECISIONS)
def test_combiner_and_meta_networks(gpu, precision):
    cfg = arch.CONFIGS["hybrid_full"]
    g = torch.Generator().manual_seed(6)
    net, orc = net_for(gpu, cfg, precision), oracle_for(cfg)
    rel = TC_LAYER_REL["bf16x3"] if precision == "bf16x3" else 2e-5
    x = (torch.randn((5, 18, 256), generator=g) * 20).float()
    got = net.engine.run_net("combiner0", x).cpu()
    ref = orc.nets["combiner0"](x.transpose(1, 2)).transpose(1, 2)
    assert (got - ref).abs().max().item() < rel * max(1.0, ref.abs().max().item())
    s = (torch.randn((4, 18, 128), generator=g) * 20).float()
    got = net.engine.run_net("meta", s).cpu().reshape(4, 3)
    ref = orc.nets["meta"](s.transpose(1, 2))
    assert (got - ref).abs().max().item() < rel * max(1.0, ref.abs().max().item())


@pytest.mark.parametrize("precision", ["bf16x3", "fp32"])
@pytest.mark.parametrize("name", ["single_tech_addendum", "hybrid_no_ensemble_addendum"])
def test_addendum_models_against_the_oracle(gpu, name, precision):
    """Transfer-learning models (build_on_top, MixtureOfExpertsAdvancedXferLearning.py:94-183): every sub-network with
    an addendum against the oracle's, and the whole forward on more sites than the golden case holds."""
    cfg = arch.CONFIGS[name]
    eng, orc = net_for(gpu, cfg, precision).engine, oracle_for(cfg)
    tol = TC_LAYER_REL["bf16x3"] if precision == "bf16x3" else 2e-5
    g = torch.Generator().manual_seed(5)
    pl = synth.make_pileups(60, coverage=9, channels=cfg.read_cin, seed=77)
    reads = pl.reads[0][:50]
    got = eng.run_net("read_convolver0", reads).cpu()
    ref = orc.nets["read_convolver0"](reads.transpose(1, 2).float()).transpose(1, 2)
    assert (got - ref).abs().max().item() < tol * max(1.0, ref.abs().max().item())
    xname = "xattn0" if not cfg.hybrid else "xattn2"
    for net_name, shape in (("compressor0", (13, 36, 64)), (xname, (13, 18, 128))):
        x = (torch.randn(shape, generator=g) * 10).float()
        got = eng.run_net(net_name, x).cpu()
        ref = orc.nets[net_name](x.transpose(1, 2))
        ref = ref.transpose(1, 2) if ref.dim() == 3 else ref.reshape(got.shape)
        assert (got - ref).abs().max().item() < tol * max(1.0, ref.abs().max().item()), net_name
    tensors, naps, nrpa, ref_seg = pl.forward_args()
    res = net_for(gpu, cfg, precision).forward(tensors, naps, nrpa, ref_seg)
    want = orc.forward(tensors, naps, nrpa, ref_seg)
    np.testing.assert_allclose(flat_result(cfg, res)[0].numpy(), flat_result(cfg, want)[0].numpy(), rtol=0,
                               atol=TOL_LOGIT[precision])


def test_deeper_addendum_runs_fused_prefix_plus_layerwise_tail(gpu):
    """Three added blocks per sub-network (the fused kernels take two): the third runs layer by layer behind the fused
    kernel, for feature networks and for the pooled expert head."""
    import dataclasses
    cfg = dataclasses.replace(arch.CONFIGS["single_tech_addendum"], name="single_tech_addendum3", addendum_blocks=3)
    params = weights.init_params(cfg, seed=13)
    from oracle import hello_oracle as O
    orc = O.OracleModel(cfg, params)
    net = gpu.MoEAttentionB200(cfg, params, device=DEV, precision="bf16x3")
    pl = synth.make_pileups(30, coverage=8, channels=cfg.read_cin, seed=31)
    tensors, naps, nrpa, ref_seg = pl.forward_args()
    before = net.engine.launch_count()
    res = net.forward(tensors, naps, nrpa, ref_seg)
    # site index, posterior, fill_meta + read convolver (1 + 2 convs + segsum) + compressor (1 + 2) + segsum + xattn (1 + 2 + pooled head)
    assert net.engine.launch_count() - before == 3 + 4 + 3 + 1 + 4
    want = orc.forward(tensors, naps, nrpa, ref_seg)
    np.testing.assert_allclose(res.reshape(-1).numpy(), want.reshape(-1).numpy(), rtol=0, atol=TOL_LOGIT["bf16x3"])
    reads = pl.reads[0][:40]
    got = net.engine.run_net("read_convolver0", reads).cpu()
    ref = orc.nets["read_convolver0"](reads.transpose(1, 2).float()).transpose(1, 2)
    assert (got - ref).abs().max().item() < TC_LAYER_REL["bf16x3"] * max(1.0, ref.abs().max().item())


# ------------------------------------------------------------------------------------- fused tcgen05 read convolver
@pytest.mark.parametrize("precision,name", [("bf16x3", "single_tech"), ("bf16x3", "single_tech_hp"),
                                            ("bf16", "single_tech")])
def test_readconv_tc_every_layer(gpu, precision, name):
    """Each of the 17 layer phases of the fused kernel against the oracle's fp32 layer outputs."""
    cfg = arch.CONFIGS[name]
    pl = synth.make_pileups(4, coverage=9, channels=cfg.read_cin, seed=21)
    reads = pl.reads[0][:31]                       # 2 full pairs of 6-read groups + a ragged tail
    eng = net_for(gpu, cfg, precision).engine
    ref = readconv_phase_reference(cfg, params_for(cfg), reads)
    for ph in range(17):
        out, dump = eng.readconv_debug(reads, ph)
        got = readconv_phase_from_dump(dump.cpu(), ph, reads.shape[0], ref[ph].shape[2])
        scale = max(1.0, ref[ph].abs().max().item())
        err = (got - ref[ph]).abs().max().item()
        assert err < TC_LAYER_REL[precision] * scale, (ph, err, scale)
    err = (out.cpu().transpose(1, 2) - ref[-1]).abs().max().item()
    assert err < TC_LAYER_REL[precision] * max(1.0, ref[-1].abs().max().item())


@pytest.mark.parametrize("n_reads", [1, 5, 6, 7, 12, 13, 151])
def test_readconv_tc_ragged_counts_and_layouts(gpu, n_reads):
    """Any number of reads (partial groups, partial group pairs, more items than one wave needs) in both layouts,
    bit-identical to each other and equal to the fp32 CUDA-core path within the bf16x3 tolerance."""
    from hello_b200 import _lib
    cfg = arch.CONFIGS["single_tech"]
    g = torch.Generator().manual_seed(n_reads)
    pl = synth.make_pileups(40, coverage=6, channels=cfg.read_cin, seed=300 + n_reads)
    reads = pl.reads[0][:n_reads]
    assert reads.shape[0] == n_reads
    tcn = net_for(gpu, cfg, "bf16x3").engine
    a, _ = tcn.readconv_debug(reads, -1, _lib.LAYOUT_RLC)
    b, _ = tcn.readconv_debug(reads.transpose(1, 2).contiguous(), -1, _lib.LAYOUT_RCL)
    assert torch.equal(a, b)
    c = tcn.run_net("read_convolver0", reads, _lib.LAYOUT_RLC)
    assert torch.equal(a, c), "run_net must route the read convolver through the same tensor-core kernel"
    ref = net_for(gpu, cfg, "fp32").engine.run_net("read_convolver0", reads, _lib.LAYOUT_RLC)
    scale = max(1.0, ref.abs().max().item())
    assert (a - ref).abs().max().item() < TC_LAYER_REL["bf16x3"] * scale


@pytest.mark.parametrize("name", ["single_tech", "single_tech_hp"])
def test_readconv_tc_unaligned_read_buffers(gpu, name):
    """The kernel stages each work item's rows with 16-byte aligned bulk copies; groups whose aligned superset would leave
    the launch's buffer (start / end of an unaligned buffer) take the global-load path.  Any base alignment, any count,
    both layouts: bit-identical to the same rows in a freshly allocated (aligned) tensor."""
    from hello_b200 import _lib
    cfg = arch.CONFIGS[name]
    pl = synth.make_pileups(60, coverage=9, channels=cfg.read_cin, seed=77)
    big = pl.reads[0].to(DEV)
    eng = net_for(gpu, cfg, "bf16x3").engine
    for k, n in ((1, 40), (2, 9), (3, 10), (5, 1), (7, 2), (4, 200), (9, big.shape[0] - 9)):
        view = big[k:k + n]
        assert view.is_contiguous()
        a, _ = eng.readconv_debug(view, -1, _lib.LAYOUT_RLC)
        b, _ = eng.readconv_debug(view.clone(), -1, _lib.LAYOUT_RLC)
        assert torch.equal(a, b), (k, n)
    rcl = big.transpose(1, 2).contiguous()
    for k, n in ((1, 13), (3, 100)):
        a, _ = eng.readconv_debug(rcl[k:k + n], -1, _lib.LAYOUT_RCL)
        b, _ = eng.readconv_debug(big[k:k + n].clone(), -1, _lib.LAYOUT_RLC)
        assert torch.equal(a, b), (k, n)


def test_readconv_tc_is_deterministic(gpu):
    cfg = arch.CONFIGS["single_tech"]
    pl = synth.make_pileups(200, coverage=10, channels=cfg.read_cin, seed=8)
    eng = net_for(gpu, cfg, "bf16x3").engine
    a, _ = eng.readconv_debug(pl.reads[0])
    b, _ = eng.readconv_debug(pl.reads[0])
    assert torch.equal(a, b)


# ------------------------------------------------------------------------------------- fused tcgen05 head networks
HEAD_CASES = [("single_tech", "compressor0", (15, 36, 64)), ("single_tech", "xattn0", (29, 18, 128)),
              ("hybrid_full", "meta", (14, 18, 128)), ("hybrid_full", "xattn2", (3, 18, 128))]


@pytest.mark.parametrize("precision", ["bf16x3", "bf16"])
@pytest.mark.parametrize("cfg_name,net_name,shape", HEAD_CASES)
def test_headconv_tc_every_layer(gpu, cfg_name, net_name, shape, precision):
    """Each of the 7 layer phases of the fused head kernel against the oracle's fp32 layer outputs, then the
    network output (features, logit, or the meta head's pre-softmax linear output)."""
    cfg = arch.CONFIGS[cfg_name]
    g = torch.Generator().manual_seed(11)
    x = (torch.randn(shape, generator=g) * 40).float()
    eng = net_for(gpu, cfg, precision).engine
    ref = head_phase_reference(cfg, params_for(cfg), net_name, x)
    for ph in range(7):
        out, dump = eng.headconv_debug(net_name, x, ph)
        got = head_phase_from_dump(dump.cpu(), ph, shape[0], shape[2], ref[ph].shape[1], ref[ph].shape[2])
        scale = max(1.0, ref[ph].abs().max().item())
        err = (got - ref[ph]).abs().max().item()
        assert err < TC_LAYER_REL[precision] * scale, (ph, err, scale)
    full = oracle_for(cfg).nets[net_name](x.transpose(1, 2))
    got = out.cpu().transpose(1, 2) if full.dim() == 3 else out.cpu().reshape(full.shape)
    scale = max(1.0, ref[-1].abs().max().item())
    assert (got - full).abs().max().item() < TC_LAYER_REL[precision] * scale


@pytest.mark.parametrize("precision", ["bf16x3", "bf16"])
@pytest.mark.parametrize("n_items", [1, 6, 7, 13, 901])
def test_combconv_tc(gpu, n_items, precision):
    """Fused tensor-core combiner (concat -> conv k3 256->512 -> conv 1x1 512->128): the 512-channel intermediate
    and the output against the oracle's fp32 layers, for partial groups and several waves of work items."""
    cfg = arch.CONFIGS["hybrid_full"]
    g = torch.Generator().manual_seed(40 + n_items)
    x = (torch.randn((n_items, 18, 256), generator=g) * 20).float()
    eng = net_for(gpu, cfg, precision).engine
    net = oracle_for(cfg).nets["combiner1"]
    trace = []
    ref = net(x.transpose(1, 2), trace)
    out, dump = eng.headconv_debug("combiner1", x, 0)
    dump = dump.cpu()
    mid = torch.stack([dump[r // 6, (r % 6) * 20:(r % 6) * 20 + 18, :].t() for r in range(n_items)])
    assert (mid - trace[0]).abs().max().item() < TC_LAYER_REL[precision] * max(1.0, trace[0].abs().max().item())
    got = out.cpu().transpose(1, 2)
    assert (got - ref).abs().max().item() < TC_LAYER_REL[precision] * max(1.0, ref.abs().max().item(), trace[0].abs().max().item())
    assert torch.equal(out, eng.run_net("combiner1", x)), "run_net routes the combiner through the same kernel"


@pytest.mark.parametrize("n_items", [1, 5, 6, 7, 12, 13, 25, 1801])
def test_headconv_tc_ragged_counts(gpu, n_items):
    """Any item count (partial groups, an empty second group, several waves of work items): run_net routes the
    head networks through the tensor-core kernel and agrees with the fp32 CUDA-core path."""
    cfg = arch.CONFIGS["single_tech"]
    g = torch.Generator().manual_seed(n_items)
    tcn, f32 = net_for(gpu, cfg, "bf16x3").engine, net_for(gpu, cfg, "fp32").engine
    for net_name, shape in (("compressor0", (n_items, 36, 64)), ("xattn0", (n_items, 18, 128))):
        x = (torch.randn(shape, generator=g) * 30).float()
        a = tcn.run_net(net_name, x)
        b, _ = tcn.headconv_debug(net_name, x)
        assert torch.equal(a, b)
        ref = f32.run_net(net_name, x)
        scale = max(1.0, ref.abs().max().item()) if net_name.startswith("compressor") else 200.0
        assert (a - ref).abs().max().item() < TC_LAYER_REL["bf16x3"] * scale, net_name
        assert torch.equal(a, tcn.run_net(net_name, x)), "deterministic"


# ------------------------------------------------------------------------------------------------ whole forward
@pytest.mark.parametrize("precision", ["fp32", "bf16x3"])
@pytest.mark.parametrize("case", GOLDEN_CASES)
def test_forward_matches_reference_golden(gpu, case, precision):
    """Same weights + same inputs as the reference run that produced tests/golden/*.npz."""
    cfg, pl, g = load_golden(case)
    assert weights.params_digest(params_for(cfg)) == str(g["digest"])
    net = net_for(gpu, cfg, precision)
    tensors, naps, nrpa, ref_seg = pl.forward_args()
    res = net.forward(tensors, naps, nrpa, ref_seg)
    logits, meta = flat_result(cfg, res)
    np.testing.assert_allclose(logits.numpy(), g["logits"], rtol=0, atol=TOL_LOGIT[precision])
    if meta is not None:
        np.testing.assert_allclose(meta.numpy(), g["meta"], rtol=0, atol=TOL_PROB[precision])
    r = net.last_result
    np.testing.assert_allclose(r.pair_prob[0].cpu().numpy(), g["pair_mixed"], rtol=0, atol=TOL_PROB[precision])
    np.testing.assert_allclose(r.pair_prob[1:].cpu().numpy(), g["pair_experts"], rtol=0, atol=TOL_PROB[precision])
    np.testing.assert_allclose(r.meta.cpu().numpy(), g["site_meta"], rtol=0, atol=TOL_PROB[precision])
    # float64 re-mix of prepareVcf.py
    mix64 = (g["pair_experts"].astype(np.float64) *
             np.repeat(g["site_meta"].astype(np.float64), np.diff(r.pair_off.numpy()), axis=0).T).sum(0)
    np.testing.assert_allclose(r.pair_mix64.cpu().numpy(), mix64, rtol=0, atol=TOL_PROB[precision])
    # genotype call: bit-exact wherever the reference's own top-2 margin exceeds the posterior tolerance
    check_calls(r, g["pair_mixed"], g["best_pair"], TOL_PROB[precision], case + "/" + precision)
    check_log_probs(r.pair_prob[0].cpu().numpy(), g["pair_mixed"], precision, case)
    check_log_probs(r.pair_prob[1:].cpu().numpy(), g["pair_experts"], precision, case)


@pytest.mark.parametrize("precision", PRECISIONS)
def test_batchnorm_built_model_matches_reference_golden(gpu, precision):
    """A model built without weight-norm (Conv1d + BatchNorm1d, eval mode): from_state_dict folds the batch-norms and the
    CUDA forward reproduces the reference's outputs (tests/golden/single_tech_batchnorm.npz, made by the reference)."""
    from helpers import batchnorm_params
    cfg, pl, g = load_golden("single_tech_batchnorm")
    state, params = batchnorm_params()
    net = gpu.MoEAttentionB200.from_state_dict(state, device=DEV, precision=precision)
    assert net.cfg.name == "single_tech"
    res = net.forward(*pl.forward_args())
    np.testing.assert_allclose(res.reshape(1, -1).cpu().numpy(), g["logits"], rtol=0, atol=TOL_LOGIT[precision])
    r = net.last_result
    np.testing.assert_allclose(r.pair_prob[0].cpu().numpy(), g["pair_mixed"], rtol=0, atol=TOL_PROB[precision])
    check_calls(r, g["pair_mixed"], g["best_pair"], TOL_PROB[precision], "single_tech_batchnorm/" + precision)
    check_log_probs(r.pair_prob[0].cpu().numpy(), g["pair_mixed"], precision, "batchnorm")


@pytest.mark.parametrize("precision", PRECISIONS)
def test_softplus_configuration_matches_reference_golden(gpu, precision):
    """moe_attention_config_single_tech_old_equivalent_layer_norm.py (Softplus in the read convolver and the expert head, no
    normalisation layers): the fused read convolver and expert head are instantiated for Softplus as well (the compressor
    keeps ReLU), the fp32 path runs layer by layer with the Softplus epilogue; both reproduce the reference's outputs
    (tests/golden/single_tech_softplus.npz)."""
    from helpers import batchnorm_params
    cfg, pl, g = load_golden("single_tech_softplus")
    state, params = batchnorm_params("single_tech_softplus")
    net = gpu.MoEAttentionB200.from_state_dict(state, softplus_nets=cfg.softplus_nets, device=DEV, precision=precision)
    assert net.cfg.name == "single_tech_softplus"
    before = net.engine.launch_count()
    res = net.forward(*pl.forward_args())
    launches = net.engine.launch_count() - before
    assert launches > 20 if precision == "fp32" else launches <= 12, launches     # fused in the tensor-core precisions
    np.testing.assert_allclose(res.reshape(1, -1).cpu().numpy(), g["logits"], rtol=0, atol=TOL_LOGIT[precision])
    r = net.last_result
    np.testing.assert_allclose(r.pair_prob[0].cpu().numpy(), g["pair_mixed"], rtol=0, atol=TOL_PROB[precision])
    check_calls(r, g["pair_mixed"], g["best_pair"], TOL_PROB[precision], "single_tech_softplus/" + precision)
    # more sites than the fixture holds, against the oracle
    pl2 = synth.make_pileups(60, coverage=14, channels=cfg.read_cin, seed=77)
    from oracle import hello_oracle as O
    want = O.OracleModel(cfg, params).forward(*pl2.forward_args())
    got = net.forward(*pl2.forward_args())
    assert (got.cpu().reshape(-1) - want.reshape(-1)).abs().max().item() < TOL_LOGIT[precision]


@pytest.mark.parametrize("case", ["legacy_hybrid_additive", "legacy_hybrid_combiners"])
@pytest.mark.parametrize("precision", PRECISIONS)
def test_legacy_hybrid_wiring_matches_reference_golden(gpu, precision, case):
    """The legacy MoEMergedAdvanced wiring with two technologies (additive hybrid features, BatchNorm-built meta): the state
    dict of the reference's own model is renamed / folded at load and the CUDA forward (HELLO_COMBINE_SUM) reproduces the
    reference's logits, meta weights, pair probabilities and calls (tests/golden/legacy_hybrid_additive.npz); with both
    ConvCombiners (BatchNorm-built, folded) it is the three-expert wiring `hybrid_full` (legacy_hybrid_combiners.npz)."""
    from helpers import legacy_params
    cfg, pl, g = load_golden(case)
    state, params = legacy_params(case)
    net = gpu.MoEAttentionB200.from_state_dict(state, device=DEV, precision=precision)
    assert net.cfg.name == cfg.name
    res = net.forward(*pl.forward_args())
    logits, meta = flat_result(cfg, res)
    np.testing.assert_allclose(logits.numpy(), g["logits"], rtol=0, atol=TOL_LOGIT[precision])
    np.testing.assert_allclose(meta.numpy(), g["meta"], rtol=0, atol=TOL_PROB[precision])
    r = net.last_result
    np.testing.assert_allclose(r.pair_prob[0].cpu().numpy(), g["pair_mixed"], rtol=0, atol=TOL_PROB[precision])
    np.testing.assert_allclose(r.pair_prob[1:].cpu().numpy(), g["pair_experts"], rtol=0, atol=TOL_PROB[precision])
    check_calls(r, g["pair_mixed"], g["best_pair"], TOL_PROB[precision], case + "/" + precision)
    check_log_probs(r.pair_prob[0].cpu().numpy(), g["pair_mixed"], precision, "legacy hybrid")
    # more sites than the fixture holds, against the oracle
    pl2 = synth.make_pileups(40, coverage=12, channels=cfg.read_cin, seed=99)
    from oracle import hello_oracle as O
    want = O.OracleModel(cfg, params).forward(*pl2.forward_args())
    got = net.forward(*pl2.forward_args())
    lg, mg = flat_result(cfg, got)
    lr, mr = flat_result(cfg, want)
    assert (lg - lr).abs().max().item() < TOL_LOGIT[precision] and (mg - mr).abs().max().item() < TOL_PROB[precision]


def check_calls(result, ref_mixed, ref_best, tol, what="", max_tight_share=MAX_TIGHT_SHARE):
    """Genotype calls bit-exact wherever the reference's own top-2 margin exceeds 2*tol; the number of tight-margin
    sites (compared only through their probabilities) is printed and bounded."""
    off = result.pair_off.numpy()
    got = result.best_pair.cpu().numpy()
    n_sites = len(off) - 1
    n_tight = n_tight_differs = 0
    for s in range(n_sites):
        probs = np.sort(ref_mixed[off[s]:off[s + 1]])[::-1]
        margin = probs[0] - probs[1] if probs.size > 1 else np.inf
        if margin > 2 * tol:
            assert tuple(got[s]) == tuple(ref_best[s]), (s, got[s], ref_best[s], probs[:3])
        else:
            n_tight += 1
            n_tight_differs += tuple(got[s]) != tuple(ref_best[s])
    print("check_calls[%s]: %d sites, %d exact-call checks, %d tight-margin sites (top-2 margin <= %.1e), %d of them "
          "called differently" % (what, n_sites, n_sites - n_tight, n_tight, 2 * tol, n_tight_differs))
    assert n_tight <= max(1, int(max_tight_share * n_sites)), (what, n_tight, n_sites)
    return n_tight


def check_log_probs(got, ref, precision, what=""):
    """Pair probabilities in log space wherever the reference value is above LOG_FLOOR."""
    got, ref = np.asarray(got, np.float64), np.asarray(ref, np.float64)
    m = ref > LOG_FLOOR
    if not m.any():
        return 0.0
    err = np.abs(np.log(np.maximum(got[m], 1e-300)) - np.log(ref[m])).max()
    assert err <= TOL_LOGP[precision], (what, precision, err)
    return float(err)


@pytest.mark.parametrize("precision", ["fp32", "bf16x3"])
@pytest.mark.parametrize("name", ["single_tech", "hybrid_ensemble2", "hybrid_full", "hybrid_no_ensemble"])
def test_forward_matches_oracle_seeded(gpu, name, precision):
    from oracle import hello_oracle as O
    cfg = arch.CONFIGS[name]
    pl = synth.make_pileups(24, coverage=14, channels=cfg.read_cin, seed=77)
    net = net_for(gpu, cfg, precision)
    res = net.forward(*pl.forward_args())
    ref = oracle_for(cfg).forward(*pl.forward_args())
    lg, mg = flat_result(cfg, res)
    lr, mr = flat_result(cfg, ref)
    assert (lg - lr).abs().max().item() < TOL_LOGIT[precision]
    if mr is not None:
        assert (mg - mr).abs().max().item() < TOL_PROB[precision]
    post = O.batched_posteriors(cfg, ref, pl.num_alleles_per_site())
    mixed = torch.cat([p[0] for p in post]).numpy()
    best = np.array([p[3] for p in post], np.int32)
    r = net.last_result
    np.testing.assert_allclose(r.pair_prob[0].cpu().numpy(), mixed, rtol=0, atol=TOL_PROB[precision])
    check_calls(r, mixed, best, TOL_PROB[precision], name + "/" + precision)
    check_log_probs(r.pair_prob[0].cpu().numpy(), mixed, precision, name)
    np.testing.assert_allclose(r.best_prob.cpu().numpy(), np.array([p[4] for p in post], np.float32), rtol=0,
                               atol=TOL_PROB[precision])


@pytest.mark.parametrize("gain", [4.0, 8.0])
def test_rescaled_weight_norm_gains(gpu, gain):
    """Random-init weights keep activations small; a trained model's weight-norm gains need not.  The gain g of the three
    stem convolutions of the read convolver is multiplied by `gain` each (every activation behind them grows by up to
    gain^3 = 64x / 512x: per-read features in the thousands, allele sums in the tens of thousands) and the pooled linear
    head's gain is divided by the same factor, so logits stay O(10) and the absolute 1e-3 bound keeps its meaning.
    bf16x3 is a relative-error scheme (hi+lo operands, fp32 accumulate), so it must hold the same tolerance."""
    from oracle import hello_oracle as O
    cfg = arch.CONFIGS["single_tech"]
    params = {k: v.clone() for k, v in params_for(cfg).items()}
    stem = [k for k, _, _ in weights.conv_keys(cfg, "read_convolver0")][:3]
    head = [k for k, _, kind in weights.conv_keys(cfg, "xattn0") if kind == "linear"]
    assert len(head) == 1
    for k in stem:
        params[k + ".weight_g"] *= gain
    params[head[0] + ".weight_g"] /= gain ** 3
    pl = synth.make_pileups(48, coverage=20, channels=cfg.read_cin, seed=314)
    orc = O.OracleModel(cfg, params)
    ref = orc.forward(*pl.forward_args())
    feat = orc.read_features(pl.reads[0].transpose(1, 2))
    base = oracle_for(cfg).read_features(pl.reads[0].transpose(1, 2))
    growth = feat.abs().max().item() / base.abs().max().item()
    assert growth > gain ** 3 / 4, growth                   # the activations really are that much larger
    net = gpu.MoEAttentionB200(cfg, params, device=DEV, precision="bf16x3")
    res = net.forward(*pl.forward_args())
    err = (res.cpu() - ref).abs().max().item()
    print("rescaled gains x%g: read features up to %.3g (%.0fx the random-init model), logits in [%.2f, %.2f], "
          "max |dlogit| = %.2e" % (gain, feat.abs().max().item(), growth, ref.min().item(), ref.max().item(), err))
    assert err < TOL_LOGIT["bf16x3"], err
    post = O.batched_posteriors(cfg, ref, pl.num_alleles_per_site())
    mixed = torch.cat([p[0] for p in post]).numpy()
    np.testing.assert_allclose(net.last_result.pair_prob[0].cpu().numpy(), mixed, rtol=0, atol=TOL_PROB["bf16x3"])
    check_calls(net.last_result, mixed, np.array([p[3] for p in post], np.int32), TOL_PROB["bf16x3"], "gain x%g" % gain)
    check_log_probs(net.last_result.pair_prob[0].cpu().numpy(), mixed, "bf16x3", "gain")


@pytest.mark.parametrize("name", ["single_tech", "hybrid_no_ensemble", "hybrid_full"])
def test_fast_mode_is_close(gpu, name):
    """bf16 (single product) is the fast mode, reported separately from the parity mode: logits within 0.15, pair
    probabilities within 2e-2, and the genotype call equal wherever the reference's top-2 margin exceeds 5e-2."""
    from oracle import hello_oracle as O
    cfg = arch.CONFIGS[name]
    pl = synth.make_pileups(60, coverage=20, channels=cfg.read_cin, seed=88)
    net = net_for(gpu, cfg, "bf16")
    res = net.forward(*pl.forward_args())
    ref = oracle_for(cfg).forward(*pl.forward_args())
    lg, _ = flat_result(cfg, res)
    lr, _ = flat_result(cfg, ref)
    assert (lg - lr).abs().max().item() < 0.15
    post = O.batched_posteriors(cfg, ref, pl.num_alleles_per_site())
    mixed = torch.cat([p[0] for p in post]).numpy()
    np.testing.assert_allclose(net.last_result.pair_prob[0].cpu().numpy(), mixed, rtol=0, atol=2e-2)
    check_calls(net.last_result, mixed, np.array([p[3] for p in post], np.int32), 2.5e-2, name + "/bf16",
                max_tight_share=0.25)        # the fast mode's loose margin (5e-2) covers ~10 % of the sites


@pytest.mark.parametrize("precision", PRECISIONS)
def test_strict_drop_in_wrapper_call(gpu, precision):
    """network(featureDict, segment) with providePredictions, as python/caller_calling.py:651-652 calls it."""
    from oracle import hello_oracle as O
    for name in ("single_tech", "hybrid_full"):
        cfg = arch.CONFIGS[name]
        pl = synth.make_pileups(4, coverage=8, channels=cfg.read_cin, seed=9)
        network = gpu.MoEMergedWrapperB200(net_for(gpu, cfg, precision)).eval()
        network.providePredictions = True
        orc = oracle_for(cfg)
        for s in range(pl.n_sites):
            fd, seg = pl.site_feature_dict(s, allele_names=["T", "AC", "A", "G"][:len(pl.site_feature_dict(s)[0])])
            with torch.no_grad():
                got = network(fd, seg)
            ref = O.wrapper_forward(orc, fd, seg, provide_predictions=True)
            assert len(got) == 5
            for dg, dr in zip(got[:4], ref[:4]):
                assert list(dg.keys()) == list(dr.keys())
                for k in dg:
                    assert dg[k].dim() == 0 and abs(float(dg[k]) - float(dr[k])) < TOL_PROB[precision]
                check_log_probs([float(dg[k]) for k in dg], [float(dr[k]) for k in dg], precision, name)
            assert (got[4] - ref[4]).abs().max().item() < TOL_PROB[precision]
            key, value, _ = O.call_genotype(ref[0])
            vals = sorted((float(v) for v in ref[0].values()), reverse=True)
            if len(vals) == 1 or vals[0] - vals[1] > 2 * TOL_PROB[precision]:
                assert network.last_call[0] == key
        network.providePredictions = False
        assert isinstance(network(fd, seg), dict)


def test_final_call_records(gpu):
    """d_call_pair / d_call_qual / d_best_expert against the oracle's final-call step (itself pinned to the
    reference's vcfRecords by tests/golden/final_calls.npz), evaluated on the kernel's own pair probabilities:
    indices bit-exact, QUAL to double rounding."""
    from oracle import hello_oracle as O
    pool = ["G", "GA", "T", "C", "GTT", "A"]
    for name in ("single_tech", "hybrid_full", "hybrid_ensemble2"):
        cfg = arch.CONFIGS[name]
        pl = synth.make_pileups(40, coverage=8, channels=cfg.read_cin, seed=123)
        naps = pl.num_alleles_per_site()
        names = [(pool[:n] if s % 2 == 0 else pool[:n][::-1]) for s, n in enumerate(naps)]
        net = net_for(gpu, cfg, "bf16x3")
        batch = gpu.DeviceBatch.from_host(pl.reads, 1, pl.allele_read_off, pl.site_allele_off, pl.ref_onehot, DEV,
                                          allele_rank=gpu.allele_ranks(names))
        r = net.engine.run(batch)
        pp, meta, off = r.pair_prob.cpu(), r.meta.cpu(), r.pair_off.numpy()
        # the `.features` records a hello_b200 caller pickles for prepareVcf (caller_calling.py:746-754)
        recs = gpu.result_feature_records(r, names, [("chr1", 100 + s, len(names[s][0])) for s in range(len(naps))])
        for s, n in enumerate(naps):
            pairs = [(names[s][i], names[s][j]) for i in range(n) for j in range(i, n)]
            preds = [{pairs[q]: pp[1 + e, off[s] + q] for q in range(len(pairs))} for e in range(3)]
            assert [list(d.items()) for d in recs[s]["expertPredictions"]] == [list(d.items()) for d in preds]
            assert np.array_equal(recs[s]["meta"], meta[s].numpy())
            ref = O.final_calls(recs[s]["expertPredictions"], recs[s]["meta"])
            got = gpu.final_calls(r, s, names[s])
            assert got["choice"] == ref["choice"]
            for key in ("expert0", "expert1", "expert2", "best", "mean"):
                assert got[key][0] == ref[key][0], (name, s, key)
                assert abs(got[key][1] - ref[key][1]) <= 1e-12 * max(1.0, abs(ref[key][1])), (name, s, key)
            key, value, qual = O.call_genotype({pairs[q]: pp[0, off[s] + q] for q in range(len(pairs))})
            assert got["mixed"][0] == key and abs(got["mixed"][1] - qual) <= 1e-12 * max(1.0, qual)
            assert tuple(r.call_pair[s, 0].tolist()) == tuple(r.best_pair[s].tolist())


# ------------------------------------------------------------------------------------------------ edge cases
@pytest.mark.parametrize("precision", PRECISIONS)
def test_chunking_does_not_change_results(gpu, precision):
    cfg = arch.CONFIGS["single_tech"]
    pl = synth.make_pileups(40, coverage=10, channels=cfg.read_cin, seed=31)
    whole = net_for(gpu, cfg, precision)
    res_a = whole.forward(*pl.forward_args())
    ra = whole.last_result
    tiny = net_for(gpu, cfg, precision, max_chunk_sites=7)
    res_b = tiny.forward(*pl.forward_args())
    rb = tiny.last_result
    assert torch.equal(res_a, res_b)
    assert torch.equal(ra.pair_prob, rb.pair_prob) and torch.equal(ra.best_pair, rb.best_pair)
    small_ws = net_for(gpu, cfg, precision, workspace_bytes=48 << 20)
    res_c = small_ws.forward(*pl.forward_args())
    assert torch.equal(res_a, res_c)


@pytest.mark.parametrize("name", ["single_tech", "hybrid_ensemble2"])
def test_forward_host_streams_ranges(gpu, name):
    """forward_host (pinned host buffers, read rows streamed range by range through hello_moe_forward_range) gives
    exactly the results of one resident forward, for any range size, and can be called repeatedly."""
    cfg = arch.CONFIGS[name]
    pl = synth.make_pileups(50, coverage=9, channels=cfg.read_cin, seed=61)
    net = net_for(gpu, cfg, "bf16x3")
    ref = net.engine.run(gpu.DeviceBatch.from_pileups(pl, DEV))
    from hello_b200 import _lib
    hb = gpu.HostBatch(pl.reads, _lib.LAYOUT_RLC, pl.allele_read_off, pl.site_allele_off, pl.ref_onehot)
    for chunk in (7, 16, 50, 1000):
        out = net.engine.forward_host(hb, chunk)                 # sync=True: results are in host memory on return
        assert out.ready()
        for got, want in zip(out.tensors(), ref.tensors()):
            assert torch.equal(got, want.cpu()), chunk
    # asynchronous form: wait() before reading; fresh buffers are not overwritten by the next call
    keep = net.engine.forward_host(hb, 16, sync=False, fresh_result=True).wait()
    again = net.engine.forward_host(hb, 16)
    assert keep.logits.data_ptr() != again.logits.data_ptr()
    for got, want in zip(keep.tensors(), ref.tensors()):
        assert torch.equal(got, want.cpu())


@pytest.mark.parametrize("precision", PRECISIONS)
def test_ragged_edges(gpu, precision):
    """One read / one allele sites, a many-allele site, an all-zero technology row."""
    from oracle import hello_oracle as O
    cfg = arch.CONFIGS["hybrid_ensemble2"]
    g = torch.Generator().manual_seed(3)
    n_alleles = [1, 6, 1, 2]
    nr0 = [1, 3, 1, 2, 1, 4, 2, 1, 5, 1]
    nr1 = [1, 1, 2, 1, 1, 1, 3, 1, 1, 2]
    r0 = torch.randint(0, 256, (sum(nr0), 150, 6), generator=g, dtype=torch.uint8)
    r1 = torch.randint(0, 256, (sum(nr1), 150, 6), generator=g, dtype=torch.uint8)
    r1[0] = 0                                              # technology without support: one all-zero row
    onehot = torch.nn.functional.one_hot(torch.randint(0, 5, (4, 150), generator=g), 5).float()
    tensors = (r0.transpose(1, 2), r1.transpose(1, 2))
    net = net_for(gpu, cfg, precision)
    res = net.forward(tensors, n_alleles, (nr0, nr1), onehot)
    ref = oracle_for(cfg).forward(tensors, n_alleles, (nr0, nr1), onehot)
    lg, mg = flat_result(cfg, res)
    lr, mr = flat_result(cfg, ref)
    scale = max(1.0, lr.abs().max().item())      # uniform random bytes: logits reach the hundreds
    assert (lg - lr).abs().max().item() < TOL_LOGIT[precision] * scale
    assert (mg - mr).abs().max().item() < TOL_PROB[precision]
    assert net.last_result.pair_prob.shape[1] == 1 + 21 + 1 + 3
    post = O.batched_posteriors(cfg, ref, n_alleles)
    mixed = torch.cat([p[0] for p in post]).numpy()
    np.testing.assert_allclose(net.last_result.pair_prob[0].cpu().numpy(), mixed, rtol=0,
                               atol=TOL_PROB[precision] * (scale if precision != "fp32" else 1.0))


def test_empty_batch_and_many_allele_site(gpu):
    """Zero sites is a valid (empty) call; a site with 40 alleles (820 genotype pairs, more than one warp pass) matches the
    oracle and its call is the argmax."""
    from oracle import hello_oracle as O
    cfg = arch.CONFIGS["single_tech"]
    net = net_for(gpu, cfg, "bf16x3")
    empty = net.forward((torch.zeros((0, 6, 150), dtype=torch.uint8), None), [], ([], None), None)
    assert empty.shape == (0, 1) and net.last_result.pair_prob.shape == (4, 0) and net.last_result.best_pair.shape == (0, 2)
    g = torch.Generator().manual_seed(12)
    n_alleles = [40, 1]
    nrpa = [1 + int(x) for x in torch.randint(0, 3, (41,), generator=g)]
    reads = torch.randint(0, 256, (sum(nrpa), 6, 150), generator=g, dtype=torch.uint8)
    res = net.forward((reads, None), n_alleles, (nrpa, None), None)
    ref = oracle_for(cfg).forward((reads, None), n_alleles, (nrpa, None), None)
    assert (res - ref).abs().max().item() < TOL_LOGIT["bf16x3"] * max(1.0, ref.abs().max().item() / 10)
    r = net.last_result
    assert r.pair_prob.shape == (4, 820 + 1)
    post = O.batched_posteriors(cfg, res.cpu(), n_alleles)          # posteriors from the kernel's own logits
    mixed = torch.cat([p[0] for p in post])
    assert (r.pair_prob[0].cpu() - mixed).abs().max().item() < 1e-5
    top = int(r.pair_prob[0, :820].argmax())
    i, j = (int(x) for x in r.best_pair[0])
    assert top == i * 40 - i * (i - 1) // 2 + (j - i)


@pytest.mark.parametrize("precision", PRECISIONS)
def test_float_inputs_and_errors(gpu, precision):
    cfg = arch.CONFIGS["single_tech"]
    pl = synth.make_pileups(3, coverage=6, channels=cfg.read_cin, seed=4)
    net = net_for(gpu, cfg, precision)
    tensors, naps, nrpa, ref = pl.forward_args()
    a = net.forward(tensors, naps, nrpa, ref)
    b = net.forward((tensors[0].float(), None), naps, nrpa, ref)       # the reference passes floats
    assert torch.equal(a, b)
    with pytest.raises(ValueError):
        net.forward((tensors[0].float() + 0.5, None), naps, nrpa, ref)
    with pytest.raises(ValueError):
        net.forward(tensors, naps, ([0] + nrpa[0][1:], None), ref)      # empty slot
    with pytest.raises(ValueError):
        net.forward(tensors, naps[:-1], nrpa, ref)                      # inconsistent CSR


@pytest.mark.parametrize("precision", PRECISIONS)
def test_tie_break_uses_allele_rank(gpu, precision):
    """Two identical alleles give exactly equal pair probabilities; the reference's sort picks the greatest key."""
    cfg = arch.CONFIGS["single_tech"]
    g = torch.Generator().manual_seed(8)
    row = torch.randint(0, 256, (3, 150, 6), generator=g, dtype=torch.uint8).float()
    network = gpu.MoEMergedWrapperB200(net_for(gpu, cfg, precision))
    for names in (["A", "C"], ["C", "A"]):
        fd = {names[0]: (row.clone(), None), names[1]: (row.clone(), None)}
        out = network(fd, torch.zeros(1, 150, 5))
        assert float(out[(names[0], names[0])]) == float(out[(names[1], names[1])])
        from oracle import hello_oracle as O
        key, _, _ = O.call_genotype(out)
        assert network.last_call[0] == key


# ------------------------------------------------------------------------------------------------ properties
@pytest.mark.parametrize("precision", PRECISIONS)
def test_read_order_and_site_independence(gpu, precision):
    """Sites are independent: a site's result does not depend on which other sites share the batch."""
    cfg = arch.CONFIGS["single_tech"]
    pl = synth.make_pileups(12, coverage=10, channels=cfg.read_cin, seed=55)
    net = net_for(gpu, cfg, precision)
    full = net.forward(*pl.forward_args())
    sao = pl.site_allele_off
    aro = pl.allele_read_off[0]
    for s in (0, 5, 11):
        a0, a1 = int(sao[s]), int(sao[s + 1])
        r0, r1 = int(aro[a0]), int(aro[a1])
        sub = net.forward((pl.reads[0][r0:r1].transpose(1, 2), None), [a1 - a0],
                          (torch.diff(aro[a0:a1 + 1]).tolist(), None), None)
        assert torch.equal(sub, full[a0:a1])


def test_properties_at_scale(gpu):
    """60 k ragged-coverage sites (15x-60x, ~2.3 M reads) through a deliberately small workspace (many chunks):
    two runs are bit-identical; 150 random sites re-run on their own give bit-identical logits, pair probabilities and
    calls (sites are independent of batch composition and chunk boundaries); a random sample agrees with the oracle;
    every site's mixture probabilities are a sub-distribution and the call is its argmax."""
    from oracle import hello_oracle as O
    cfg = arch.CONFIGS["single_tech"]
    pl = synth.make_pileups(60_000, coverage=(15, 60), channels=cfg.read_cin, seed=2024, device=DEV)
    net = net_for(gpu, cfg, "bf16x3", workspace_bytes=1 << 30)
    batch = gpu.DeviceBatch.from_pileups(pl, DEV)
    a = net.engine.run(batch)
    logits_a, pp_a, bp_a = a.logits.clone(), a.pair_prob.clone(), a.best_pair.clone()
    b = net.engine.run(batch)
    assert torch.equal(logits_a, b.logits) and torch.equal(pp_a, b.pair_prob) and torch.equal(bp_a, b.best_pair)
    sao, aro, off = pl.site_allele_off, pl.allele_read_off[0], a.pair_off
    rng = np.random.default_rng(3)
    picks = np.sort(rng.choice(60_000, 150, replace=False))
    reads = torch.cat([pl.reads[0][int(aro[sao[s]]):int(aro[sao[s + 1]])] for s in picks])
    nrpa = torch.cat([torch.diff(aro[int(sao[s]):int(sao[s + 1]) + 1]) for s in picks]).tolist()
    naps = [int(sao[s + 1] - sao[s]) for s in picks]
    sub_net = net_for(gpu, cfg, "bf16x3")
    sub = sub_net.forward((reads.transpose(1, 2), None), naps, (nrpa, None), None)
    r = sub_net.last_result
    idx_a = torch.cat([torch.arange(int(sao[s]), int(sao[s + 1])) for s in picks])
    idx_p = torch.cat([torch.arange(int(off[s]), int(off[s + 1])) for s in picks])
    assert torch.equal(sub.reshape(-1).cpu(), logits_a[0].cpu()[idx_a])
    assert torch.equal(r.pair_prob.cpu(), pp_a.cpu()[:, idx_p]) and torch.equal(r.best_pair.cpu(), bp_a.cpu()[picks])
    # oracle on 40 of them
    orc = oracle_for(cfg)
    some = picks[:40]
    reads_s = torch.cat([pl.reads[0][int(aro[sao[s]]):int(aro[sao[s + 1]])] for s in some]).cpu()
    nrpa_s = torch.cat([torch.diff(aro[int(sao[s]):int(sao[s + 1]) + 1]) for s in some]).tolist()
    ref = orc.forward((reads_s.transpose(1, 2), None), [int(sao[s + 1] - sao[s]) for s in some], (nrpa_s, None), None)
    idx_s = torch.cat([torch.arange(int(sao[s]), int(sao[s + 1])) for s in some])
    assert (logits_a[0].cpu()[idx_s] - ref.reshape(-1)).abs().max().item() < TOL_LOGIT["bf16x3"]
    # per-site sanity over the whole batch
    pp = pp_a[0].cpu()
    site_of_pair = torch.repeat_interleave(torch.arange(60_000), torch.diff(off))
    mass = torch.zeros(60_000).index_add_(0, site_of_pair, pp)
    assert float(mass.max()) <= 1.0 + 1e-4 and float(pp.min()) >= 0.0
    best = torch.zeros(60_000).index_reduce_(0, site_of_pair, pp, "amax", include_self=False)
    assert torch.equal(best, a.best_prob.cpu())


@pytest.mark.parametrize("name", ["hybrid_no_ensemble", "hybrid_full"])
def test_hybrid_site_independence_across_chunks(gpu, name):
    """Two technologies, combiners and (hybrid_full) three experts + meta gate: 3000 sites through a small workspace
    (many chunks) against the same sites re-run 40 at a time -- bit-identical logits, meta weights, pair probabilities
    and calls, i.e. no kernel's result depends on which items share its tiles."""
    cfg = arch.CONFIGS[name]
    pl = synth.make_pileups(3000, coverage=(8, 30), channels=cfg.read_cin, seed=777, device=DEV)
    net = net_for(gpu, cfg, "bf16x3", workspace_bytes=256 << 20)
    full = net.engine.run(gpu.DeviceBatch.from_pileups(pl, DEV))
    full = [t.clone() for t in (full.logits, full.meta, full.pair_prob, full.best_pair, full.call_pair)]
    off = net.last_result.pair_off if net.last_result is not None else None
    sub_net = net_for(gpu, cfg, "bf16x3")
    sao = pl.site_allele_off
    from hello_b200 import _lib
    for s0 in (0, 1234, 2960):
        s1 = s0 + 40
        a0, a1 = int(sao[s0]), int(sao[s1])
        reads, offs = [], []
        for t in range(2):
            aro = pl.allele_read_off[t]
            r0, r1 = int(aro[a0]), int(aro[a1])
            reads.append(pl.reads[t][r0:r1])
            offs.append(aro[a0:a1 + 1] - r0)
        b = gpu.DeviceBatch.from_host(reads, _lib.LAYOUT_RLC, offs, sao[s0:s1 + 1] - a0, pl.ref_onehot[s0:s1], DEV)
        r = sub_net.engine.run(b)
        p0 = int(synth.pair_offsets(sao)[s0]); p1 = int(synth.pair_offsets(sao)[s1])
        assert torch.equal(r.logits, full[0][:, a0:a1]) and torch.equal(r.meta, full[1][s0:s1])
        assert torch.equal(r.pair_prob, full[2][:, p0:p1]) and torch.equal(r.best_pair, full[3][s0:s1])
        assert torch.equal(r.call_pair, full[4][s0:s1])


def test_huge_allele_and_partition_invariance(gpu):
    """An allele supported by 4000 reads next to alleles with 1-2 reads: the fused reads->alleles sum must add every
    allele's reads in read order whatever the CTA work partition is, so the site's results are bit-identical whether it
    is scored alone or surrounded by other sites (which moves every partition boundary), and they match the oracle."""
    cfg = arch.CONFIGS["single_tech"]
    g = torch.Generator().manual_seed(4000)
    n_alleles = [3]
    nrpa = [4000, 1, 2]
    reads = torch.randint(0, 256, (sum(nrpa), 6, 150), generator=g, dtype=torch.uint8)
    reads[:, 2:] = reads[:, 2:] // 4                       # keep the 4000-read sum in a sane range
    net = net_for(gpu, cfg, "bf16x3")
    alone = net.forward((reads, None), n_alleles, (nrpa, None), None).clone()
    pp_alone = net.last_result.pair_prob.clone()
    pl = synth.make_pileups(300, coverage=11, channels=cfg.read_cin, seed=9)
    t, naps, (nr0, _), _ = pl.forward_args()
    for k in (0, 150, 300):                                # the big site first, in the middle, last
        sao = pl.site_allele_off
        a_k = int(sao[k])
        r_k = int(pl.allele_read_off[0][a_k])
        mixed_reads = torch.cat([t[0][:r_k], reads, t[0][r_k:]])
        mixed = net.forward((mixed_reads, None), naps[:k] + n_alleles + naps[k:], (nr0[:a_k] + nrpa + nr0[a_k:], None), None)
        assert torch.equal(mixed[a_k:a_k + 3], alone), k
        p_k = int(net.last_result.pair_off[k])
        assert torch.equal(net.last_result.pair_prob[:, p_k:p_k + 6], pp_alone), k
    ref = oracle_for(cfg).forward((reads, None), n_alleles, (nrpa, None), None)
    assert (alone - ref).abs().max().item() < 1e-3 * max(1.0, ref.abs().max().item())


def test_c_abi_status_codes(gpu):
    """include/hello_moe.h: every entry point answers with a negative hello_status and a message, never an exception or
    a crash, and a failed call leaves the handle usable."""
    import ctypes as C
    from hello_b200 import _lib
    lib = _lib.load()
    cfg = arch.CONFIGS["single_tech"]
    blob = weights.pack_blob(cfg, params_for(cfg))

    def make_cfg(**over):
        c = _lib.HelloCfg()
        c.struct_size = C.sizeof(_lib.HelloCfg)
        c.n_tech = 1
        c.read_channels[0] = 6
        c.xattn_present[0] = 1
        c.meta_kind = _lib.META_NONE
        c.feature_length = arch.FEATURE_LENGTH
        c.precision = _lib.PRECISIONS["bf16x3"]
        for k, v in over.items():
            setattr(c, k, v)
        return c

    def create(blob_bytes, c):
        h = C.c_void_p()
        buf = (C.c_char * len(blob_bytes)).from_buffer_copy(blob_bytes)
        rc = lib.hello_moe_create(buf, len(blob_bytes), C.byref(c), 0, C.byref(h))
        return rc, h, lib.hello_moe_last_error(None).decode()

    rc, h, msg = create(blob, make_cfg(struct_size=8))
    assert rc == -1 and not h.value and "struct_size" in msg                      # HELLO_ERR_ARG
    rc, h, msg = create(b"NOTABLOB" + blob[8:], make_cfg())
    assert rc == -2 and not h.value and "magic" in msg                            # HELLO_ERR_BLOB
    rc, h, msg = create(blob[:len(blob) // 2], make_cfg())
    assert rc == -2 and not h.value and "truncated" in msg
    rc, h, msg = create(blob, make_cfg(n_tech=2))
    assert rc == -5 and not h.value and "wiring" in msg                           # HELLO_ERR_UNSUPPORTED
    c7 = make_cfg()
    c7.read_channels[0] = 7
    rc, h, msg = create(blob, c7)
    assert rc == -5 and not h.value

    rc, h, msg = create(blob, make_cfg())
    assert rc == 0 and h.value
    try:
        assert lib.hello_moe_forward(h, None, None, None, 0, None) == -1          # null batch / result
        eng = net_for(gpu, cfg, "bf16x3").engine
        pl = synth.make_pileups(50, coverage=8, channels=cfg.read_cin, seed=3)
        batch = gpu.DeviceBatch.from_pileups(pl, DEV)
        good = eng.run(batch)
        tiny = torch.empty(4096, dtype=torch.uint8, device=DEV)                   # not even one site fits
        with pytest.raises(_lib.HelloMoEError) as ei:
            eng.run(batch, workspace=tiny)
        assert "(-4)" in str(ei.value) or "workspace" in str(ei.value).lower()    # HELLO_ERR_WORKSPACE
        aro = pl.allele_read_off[0].clone()
        aro[3] = aro[2]                                                           # an allele without rows
        empty = gpu.DeviceBatch.from_host(pl.reads, _lib.LAYOUT_RLC, (aro,), pl.site_allele_off, None, DEV)
        with pytest.raises(_lib.HelloMoEError) as ei:
            eng.run(empty)
        assert "(-1)" in str(ei.value) and "at least one read row" in str(ei.value)   # HELLO_ERR_ARG
        again = eng.run(batch)                                                    # the handle is still good
        assert torch.equal(good.logits, again.logits) and torch.equal(good.best_pair, again.best_pair)
    finally:
        lib.hello_moe_destroy(h)
    lib.hello_moe_destroy(None)                                                   # documented as a no-op
