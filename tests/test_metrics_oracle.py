"""CPU: the numpy restatement of clustering.py's metrics against scipy (the library the reference calls)."""
import numpy as np
from scipy.spatial.distance import pdist, squareform

from oracle import metrics_oracle as MO


def test_pdist_restatement_matches_scipy():
    x = np.random.RandomState(0).randn(11, 3)
    assert np.allclose(MO.pdist(x), pdist(x), rtol=1e-14, atol=0)


def test_dmae_known_answer():
    x = np.array([[0.0, 0, 0], [1, 0, 0], [0, 2, 0]])
    y = np.array([[0.0, 0, 0], [2, 0, 0], [0, 2, 0]])
    dm = lambda p: squareform(pdist(p))  # noqa: E731
    # pairs: |1-2| + |2-2| + |sqrt5 - sqrt8| over 3 pairs
    want = (1.0 + 0.0 + abs(5 ** 0.5 - 8 ** 0.5)) / 3
    assert abs(MO.calc_DMAE(dm(x), dm(y)) - want) < 1e-15


def test_minimum_matches_recovers_a_permutation():
    rng = np.random.RandomState(1)
    ref = rng.randn(6, 3)
    perm = [0, 2, 1, 3, 5, 4]
    prb = ref[np.argsort(perm)]  # prb[perm] == ref
    matches = [[0, 1, 2, 3, 4, 5], perm, [0, 2, 1, 3, 4, 5]]
    assert MO.get_minimum_matches(ref, prb, matches, return_type="match") == perm
    assert MO.get_minimum_matches(ref, prb, matches) < 1e-25
