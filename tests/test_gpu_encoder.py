"""GPU read-feature encoder (include/hello_encode.h) against the reference-made golden encodings and the oracle
restatement of the C++: bit-exact bytes."""
import os

import numpy as np
import pytest
import torch

from hello_b200 import arch
from helpers import GOLDEN, params_for

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


@pytest.fixture(scope="module")
def encoder():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from hello_b200 import encoder
    return encoder


def test_encoder_matches_reference_specification(encoder):
    """The reference's fixture pileups and seeded reads, encodings made by python/test_aligner.py:create_read_encoding."""
    from test_encoder_oracle import golden_cases
    n = 0
    for k, site, L, with_hp, want in golden_cases():
        got = encoder.SiteEncoderB200(site, DEV).computeFeaturesColoredSimple("x", L, False, with_hp)
        assert got.shape == (1, L, 7 if with_hp else 6) and got.dtype == np.uint8
        assert np.array_equal(got[0], want), k
        n += 1
    assert n >= 60


def test_encoder_matches_compiled_reference_encoder(encoder):
    """tests/golden/encoder_cpp.npz: outputs of the reference's own compiled C++ computeFeaturesColoredSimple on the
    corner cases its Python specification cannot express (clips, border indels, varying qualities, N bases)."""
    from test_encoder_oracle import cpp_golden_cases
    n, encs = 0, {}
    for site, allele, L, pac, hp, want in cpp_golden_cases():
        enc = encs.get(id(site))
        if enc is None:
            enc = encs[id(site)] = encoder.SiteEncoderB200(site, DEV)
        got = enc.computeFeaturesColoredSimple(allele, L, pac, hp)
        assert got.shape == want.shape and np.array_equal(got, want), (allele, L, pac, hp)
        n += 1
    assert n >= 400


@pytest.mark.parametrize("long_reads", [False, True])
def test_encoder_matches_oracle_everywhere(encoder, long_reads):
    """Window borders, clips, skips, 'N' bases, long insertions, varying qualities, both technologies, 6 and 7
    channels, unsupported alleles: per-call results equal the oracle's for every (allele, technology, hp) call."""
    from oracle import encoder_oracle as E
    rng = np.random.default_rng(7 + long_reads)
    for trial in range(12):
        site = E.random_site(rng, n_reads=14, long_reads=long_reads, border_cases=True)
        if trial % 3 == 0:                                   # mixed technologies at one site
            site.pacbio = [bool(i % 2) for i in range(len(site.reads))]
        enc = encoder.SiteEncoderB200(site, DEV)
        for allele in list(site.supports) + ["absent"]:
            for pacbio in (False, True):
                for hp in (False, True):
                    want = E.compute_features_colored_simple(site, allele, 150, pacbio, hp)
                    got = enc.computeFeaturesColoredSimple(allele, 150, pacbio, hp)
                    assert got.shape == want.shape and np.array_equal(got, want), (trial, allele, pacbio, hp)
    site = E.random_site(rng, n_reads=5, long_reads=long_reads)
    for L in (10, 33, 149, 160):                              # other feature lengths (the reference's test uses 10)
        want = E.compute_features_colored_simple(site, "ref", L, long_reads, False)
        got = encoder.SiteEncoderB200(site, DEV).computeFeaturesColoredSimple("ref", L, long_reads, False)
        assert np.array_equal(got, want), L


def test_batched_encoding_feeds_the_network(encoder):
    """encode_sites() -> DeviceBatch -> forward equals the forward on the oracle-encoded pileups; the batched rows are
    the per-call arrays concatenated in network order; encoding is deterministic."""
    from hello_b200 import model, _lib
    from oracle import encoder_oracle as E
    rng = np.random.default_rng(99)
    sites = [E.random_site(rng, n_reads=int(rng.integers(3, 20)), border_cases=False) for _ in range(25)]
    alleles = [[a for a in s.supports if a != "unsupported"] + (["unsupported"] if i % 5 == 0 else []) for i, s in enumerate(sites)]
    reads, offs, sao = encoder.encode_sites(sites, alleles, ((False, 6),), DEV)
    want = np.concatenate([E.compute_features_colored_simple(s, a, 150, False, False) for s, al in zip(sites, alleles) for a in al])
    assert np.array_equal(reads[0].cpu().numpy(), want)
    again, _, _ = encoder.encode_sites(sites, alleles, ((False, 6),), DEV)
    assert torch.equal(reads[0], again[0])
    cfg = arch.CONFIGS["single_tech"]
    net = model.MoEAttentionB200(cfg, params_for(cfg), device=DEV, precision="bf16x3")
    got = net.engine.run(model.DeviceBatch.from_host(reads, _lib.LAYOUT_RLC, offs, sao, None, DEV))
    ref = net.engine.run(model.DeviceBatch.from_host((torch.from_numpy(want),), _lib.LAYOUT_RLC, offs, sao, None, DEV))
    assert torch.equal(got.logits, ref.logits) and torch.equal(got.best_pair, ref.best_pair)
    assert int(offs[0][-1]) == want.shape[0] and int(sao[-1]) == sum(len(a) for a in alleles)


def test_encoder_rejects_bad_arguments(encoder):
    import ctypes as C
    lib = encoder._load()
    b = encoder.HelloEncodeBatch()
    b.n_rows, b.feature_length, b.channels = 4, 150, 5
    assert lib.hello_encode_reads(C.byref(b), 1, None) == -1 and b"channels" in lib.hello_encode_last_error()
    b.channels = 6
    assert lib.hello_encode_reads(C.byref(b), 1, None) == -1 and b"missing" in lib.hello_encode_last_error()
    b.n_rows = 0
    assert lib.hello_encode_reads(C.byref(b), 1, None) == 0


def test_encoder_at_scale_properties(encoder):
    """1.5 M rows (50 k sites x 30x): a random sample of rows equals the oracle; encoding a permutation of the rows
    gives the permuted bytes (rows are independent); a second launch is bit-identical."""
    from hello_b200 import synth
    from oracle import encoder_oracle as E
    packed, row_read, row_site = synth.make_packed_reads(50_000, 30, seed=5)
    dp = encoder.DevicePackedReads(packed, DEV)
    out = dp.encode(row_read, row_site, 6)
    R = row_read.size
    assert out.shape == (R, 150, 6)
    rng = np.random.default_rng(1)
    perm = rng.permutation(R).astype(np.int64)
    out_p = dp.encode(row_read[perm], row_site[perm], 6)
    assert torch.equal(out_p, out[torch.from_numpy(perm).to(DEV)])
    assert torch.equal(dp.encode(row_read, row_site, 6), out)
    host = out.cpu().numpy()
    p = packed
    for r in rng.choice(R, 300, replace=False):
        s = int(row_site[r])
        site = E.SitePileup([bytes(p.bases[p.read_off[r]:p.read_off[r + 1]]).decode()],
                            [list(p.quals[p.read_off[r]:p.read_off[r + 1]])],
                            [[(int(c) & 15, int(c) >> 4) for c in p.cigars[p.cigar_off[r]:p.cigar_off[r + 1]]]],
                            [int(p.ref_start[r])], [int(p.mapq[r])], [int(p.orientation[r])], [False], [0],
                            bytes(p.reference[p.ref_off[s]:p.ref_off[s + 1]]).decode(), int(p.window_start[s]),
                            int(p.assembly_start[s]), int(p.assembly_stop[s]), {"a": [0]})
        assert np.array_equal(host[r], E.compute_features_colored_simple(site, "a", 150, False, False)[0]), r


@pytest.mark.parametrize("hp", [False, True])
def test_forward_host_packed_equals_encode_then_forward(encoder, hp):
    """MoEEngine.forward_host_packed (aligned reads on the host -> H2D of the packed reads per range -> hello_encode_reads ->
    hello_moe_forward_range) gives bit-identical results to encoding the whole batch once and scoring the resident rows, for
    any range size, including the all-zero row of an allele without support."""
    from hello_b200 import model, synth
    cfg = arch.CONFIGS["single_tech_hp" if hp else "single_tech"]
    packed, rr, rs = synth.make_packed_reads(300, coverage=12, seed=21, hp=hp)
    aro, sao = synth.packed_allele_csr(np.diff(packed.read_base), seed=21)
    rr = rr.copy()
    rr[int(aro[3])] = -1 if int(aro[4]) - int(aro[3]) == 1 else rr[int(aro[3])]      # an allele without support, if single-row
    eng = model.MoEEngine(cfg, params_for(cfg), device=DEV, precision="bf16x3")
    rows = encoder.DevicePackedReads(packed, DEV).encode(rr, rs, cfg.read_cin[0])
    from hello_b200 import _lib
    want = eng.run(model.DeviceBatch.from_host((rows,), _lib.LAYOUT_RLC, (aro,), sao, None, DEV))
    hpb = model.HostPackedBatch(packed, (rr,), (rs,), (aro,), sao)
    assert hpb.input_nbytes() < 0.5 * rows.numel()                   # ~370 B per row cross PCIe instead of 900 / 1050
    for chunk in (37, 128, 1000):
        got = eng.forward_host_packed(hpb, chunk)
        for g, w in zip(got.tensors(), want.tensors()):
            assert torch.equal(g, w.cpu()), chunk
