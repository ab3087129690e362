"""GPU: post-sampling geometry metrics against the numpy restatement of clustering.py:98-135.
fp64 on both sides; bound 1e-12 relative (different summation order only), arg-min index exact."""
import itertools

import numpy as np
import pytest
import torch

from oracle import metrics_oracle as MO
from tsdiff_b200 import metrics as M

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _geoms(n, batch, seed):
    rng = np.random.RandomState(seed)
    ref = rng.randn(n, 3) * 2.0
    prb = ref[None] + 0.3 * rng.randn(batch, n, 3)
    return ref, prb


@pytest.mark.parametrize("n,mape", [(2, False), (13, False), (13, True), (60, False)])
def test_calc_dmae_vs_numpy(n, mape):
    ref, prb = _geoms(n, 7, n)
    dm = lambda x: np.sqrt(((x[:, None] - x[None]) ** 2).sum(-1))  # noqa: E731
    want = np.array([MO.calc_DMAE(dm(ref), dm(g), mape=mape) for g in prb])
    dm_ref = torch.tensor(dm(ref), device=DEV)
    dm_g = torch.tensor(np.stack([dm(g) for g in prb]), device=DEV)
    got = M.calc_DMAE(dm_ref, dm_g, mape=mape).cpu().numpy()
    assert np.allclose(got, want, rtol=1e-12, atol=0)
    one = M.calc_DMAE(dm_ref, dm_g[3], mape=mape)
    assert one.dim() == 0 and abs(float(one) - want[3]) <= 1e-12 * abs(want[3])
    got_p = M.calc_DMAE_from_positions(torch.tensor(ref, device=DEV), torch.tensor(prb, device=DEV), mape=mape).cpu().numpy()
    assert np.allclose(got_p, want, rtol=1e-12, atol=0)


def test_get_minimum_matches_vs_numpy():
    n = 9
    ref, prb = _geoms(n, 5, 3)
    # permutations of three equivalent hydrogens x two equivalent carbons (what index_align enumerates)
    matches = []
    for ph in itertools.permutations([6, 7, 8]):
        for pc in itertools.permutations([1, 2]):
            m = list(range(n))
            m[6:9] = ph
            m[1:3] = pc
            matches.append(m)
    rng = np.random.RandomState(0)
    prb = np.stack([g[matches[rng.randint(len(matches))]] for g in prb])  # scramble: the minimum is not the identity
    want_v = np.array([MO.get_minimum_matches(ref, g, matches) for g in prb])
    want_m = [MO.get_minimum_matches(ref, g, matches, return_type="match") for g in prb]
    t_ref, t_prb = torch.tensor(ref, device=DEV), torch.tensor(prb, device=DEV)
    got_v = M.get_minimum_matches(t_ref, t_prb, matches).cpu().numpy()
    assert np.allclose(got_v, want_v, rtol=1e-12, atol=1e-300)
    got_m = M.get_minimum_matches(t_ref, t_prb, matches, return_type="match").cpu().tolist()
    assert got_m == [list(m) for m in want_m]
    v1 = M.get_minimum_matches(t_ref, t_prb[2], matches)
    assert v1.dim() == 0 and abs(float(v1) - want_v[2]) <= 1e-12 * max(want_v[2], 1e-30)


def test_get_minimum_matches_many_permutations_and_ties():
    """More permutations than one CTA covers (5040 > 256) and an exact tie: identical atoms give equal metric values,
    the FIRST minimising permutation wins (list.index(min(...)) in the reference)."""
    n = 8
    ref, prb = _geoms(n, 3, 11)
    matches = [list(p) + [7] for p in itertools.permutations(range(7))]
    want_v = np.array([MO.get_minimum_matches(ref, g, matches) for g in prb])
    t_ref, t_prb = torch.tensor(ref, device=DEV), torch.tensor(prb, device=DEV)
    got_v = M.get_minimum_matches(t_ref, t_prb, matches).cpu().numpy()
    assert np.allclose(got_v, want_v, rtol=1e-12, atol=1e-300)
    prb_tie = prb.copy()
    prb_tie[:, 1] = prb_tie[:, 0]  # atoms 0 and 1 coincide: swapping them leaves the metric unchanged
    want_m = [MO.get_minimum_matches(ref, g, matches, return_type="match") for g in prb_tie]
    got_m = M.get_minimum_matches(t_ref, torch.tensor(prb_tie, device=DEV), matches, return_type="match").cpu().tolist()
    assert got_m == [list(m) for m in want_m]


def test_metrics_refuse_cpu_tensors():
    from tsdiff_b200._lib import TsdError
    with pytest.raises(TsdError):
        M.calc_DMAE(torch.zeros(3, 3), torch.zeros(3, 3))
    with pytest.raises(ValueError):
        M.get_minimum_matches(torch.zeros(3, 3, device=DEV), torch.zeros(3, 3, device=DEV), [])
