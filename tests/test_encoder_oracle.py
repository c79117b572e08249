"""The encoder oracle (restatement of the C++ computeFeaturesColoredSimple) against encodings made by the reference's own
Python specification (tests/golden/encoder.npz, oracle/gen_encoder_golden.py)."""
import os

import numpy as np

from helpers import GOLDEN
from oracle import encoder_oracle as E


def golden_cases():
    g = np.load(os.path.join(GOLDEN, "encoder.npz"), allow_pickle=True)
    for k in range(len(g["reads"])):
        a0, a1, L, rs, mq, ori, hp = (int(x) for x in g["meta"][k])
        site = E.SitePileup([str(g["reads"][k])], [list(g["quals"][k])], [[tuple(c) for c in g["cigars"][k]]], [rs], [mq], [ori],
                            [False], [max(hp, 0)], str(g["references"][k]), 0, a0, a1, {"x": [0]})
        yield k, site, L, hp >= 0, g["encodings"][k]


def test_oracle_matches_reference_specification():
    n = 0
    for k, site, L, with_hp, want in golden_cases():
        got = E.compute_features_colored_simple(site, "x", L, False, with_hp)
        assert got.shape == (1, L, 7 if with_hp else 6)
        assert np.array_equal(got[0], want), k
        n += 1
    assert n >= 60


def test_oracle_edge_semantics():
    """Restated-from-C++ corners: no support -> one zero row; technology filter; deletion whose preceding base is left
    of the window draws nothing; insertion takes the minimum quality."""
    ref = "ACGT" * 100
    site = E.SitePileup(["ACGTACGTAC", "ACGTACGTAC"], [[30] * 10, [10, 40, 40, 5, 40, 40, 40, 40, 40, 40]],
                        [[(0, 4), (2, 3), (0, 6)], [(0, 2), (1, 3), (0, 5)]], [121, 130], [60, 20], [1, -1], [False, True],
                        [0, 2], ref, 0, 200, 201, {"a": [0, 1], "b": []})
    assert E.compute_features_colored_simple(site, "b", 150, False, False).shape == (1, 150, 6)
    assert not E.compute_features_colored_simple(site, "b", 150, False, False).any()
    ill = E.compute_features_colored_simple(site, "a", 150, False, False)
    pac = E.compute_features_colored_simple(site, "a", 150, True, True)
    assert ill.shape == (1, 150, 6) and pac.shape == (1, 150, 7)
    # window = [125, 275): read 0 starts at 121, its deletion sits at 125..127 with the preceding base 124 outside
    assert not ill[0, 0:3].any() and ill[0, 3, E.READ_BASE] != 0
    # read 1: insertion of 3 bases after position 131 -> feature 6 carries '*' and min(qual[1:5]) = 5
    assert pac[0, 6, E.READ_BASE] == 0 and pac[0, 6, E.READ_QUAL] == E.base_quality_color(5) and pac[0, 6, E.HP] == 240


def cpp_golden_cases():
    """tests/golden/encoder_cpp.npz: whole sites + the outputs of the reference's COMPILED C++ encoder
    (oracle/gen_encoder_cpp_golden.py, AlleleSearcherLiteFiltered.cpp:1031-1180 built by oracle/Makefile)."""
    g = np.load(os.path.join(GOLDEN, "encoder_cpp.npz"))
    sites = []
    for s in range(int(g["n_sites"])):
        p = "s%d_" % s
        reads = [str(r) for r in g[p + "reads"]]
        qoff = np.concatenate([[0], np.cumsum([len(r) for r in reads])])
        quals = [[int(x) for x in g[p + "quals"][qoff[i]:qoff[i + 1]]] for i in range(len(reads))]
        coff = np.concatenate([[0], np.cumsum(g[p + "cig_n"])])
        cig = [[(int(a), int(b)) for a, b in g[p + "cig"][coff[i]:coff[i + 1]]] for i in range(len(reads))]
        pr = g[p + "per_read"]
        soff = np.concatenate([[0], np.cumsum(g[p + "sup_n"])])
        supports = {str(a): [int(x) for x in g[p + "sup"][soff[i]:soff[i + 1]]] for i, a in enumerate(g[p + "alleles"])}
        w, a0, a1 = (int(x) for x in g[p + "loc"])
        sites.append(E.SitePileup(reads, quals, cig, [int(x) for x in pr[0]], [int(x) for x in pr[1]], [int(x) for x in pr[2]],
                                  [bool(x) for x in pr[3]], [int(x) for x in pr[4]], str(g[p + "ref"]), w, a0, a1, supports))
    pos = 0
    for (s, ai, pac, hp, L), rows in zip(g["queries"], g["out_rows"]):
        n = int(rows) * int(L) * (7 if hp else 6)
        want = g["out"][pos:pos + n].reshape(int(rows), int(L), 7 if hp else 6)
        pos += n
        yield sites[int(s)], list(sites[int(s)].supports)[int(ai)], int(L), bool(pac), bool(hp), want
    assert pos == g["out"].size


def test_oracle_matches_compiled_reference_encoder():
    """Every corner the Python specification cannot pin (soft / hard clips, indels at the window borders and at the start
    of a read, insertions into reads of varying quality, N bases, technology filter, no-support row): bit-exact against
    the reference's own C++."""
    n = rows = 0
    for site, allele, L, pac, hp, want in cpp_golden_cases():
        got = E.compute_features_colored_simple(site, allele, L, pac, hp)
        assert got.shape == want.shape and np.array_equal(got, want), (allele, L, pac, hp)
        n += 1
        rows += want.shape[0]
    assert n >= 400 and rows >= 800


def test_oracle_matches_compiled_reference_encoder_live():
    """With oracle/_ref/libref_encoder.so present (built by __graft_entry__.build() where /root/reference exists): fresh
    random sites through both."""
    import pytest
    from oracle import ref_encoder as R
    if not R.available():
        pytest.skip("oracle/_ref/libref_encoder.so not built (needs /root/reference)")
    rng = np.random.default_rng(424242)
    for k in range(60):
        long_reads = k % 3 == 2
        site = E.random_site(rng, n_reads=9, long_reads=long_reads, border_cases=True, window_start=500 if k % 2 else 0)
        for allele in site.supports:
            for hp in (False, True):
                a = E.compute_features_colored_simple(site, allele, 150, long_reads, hp)
                b = R.compute_features_colored_simple(site, allele, 150, long_reads, hp)
                assert a.shape == b.shape and np.array_equal(a, b), (k, allele, hp)
