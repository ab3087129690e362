"""TEST INFRASTRUCTURE ONLY: numpy restatement of the reference's geometry metrics.
clustering.py:98-105 (`calc_DMAE`) and :123-135 (`get_minimum_matches`); scipy.spatial.distance.pdist is restated
as the i < j row-major list of Euclidean distances."""
import numpy as np


def pdist(x):
    x = np.asarray(x, dtype=np.float64)
    i, j = np.triu_indices(len(x), k=1)
    return np.sqrt(((x[i] - x[j]) ** 2).sum(-1))


def calc_DMAE(dm_ref, dm_guess, mape=False):
    if mape:
        retval = abs(dm_ref - dm_guess) / dm_ref
    else:
        retval = abs(dm_ref - dm_guess)
    return np.triu(retval, k=1).sum() / len(dm_ref) / (len(dm_ref) - 1) * 2


def get_minimum_matches(ref, prb, matches, return_type="value"):
    d_ref = pdist(ref)
    bins = [((d_ref - pdist(np.asarray(prb)[list(m)])) ** 2).sum() for m in matches]
    if return_type == "value":
        return min(bins)
    return matches[bins.index(min(bins))]
