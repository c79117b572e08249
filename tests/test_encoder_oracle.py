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
