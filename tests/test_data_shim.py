"""Host-side tests of the sampling.py caller shim (tsdiff_b200/data.py): collation semantics of
Batch.from_data_list / to_data_list, batching (sampling.py:24-42) and the pickle round trip."""
import os

import pytest
import torch

from tsdiff_b200.data import Batch, Data, batching, count_nodes_per_graph, load_samples, save_samples
from tsdiff_b200.synthetic import make_batch

from conftest import GOLDEN


def _reactions(sizes, seed=5):
    g = make_batch(len(sizes), seed=seed, sizes=sizes)
    out, off = [], 0
    for k, n in enumerate(sizes):
        sel = (g["bond_index"][0] >= off) & (g["bond_index"][0] < off + n)
        out.append(count_nodes_per_graph(Data(
            atom_type=g["atom_type"][off:off + n], r_feat=g["r_feat"][off:off + n], p_feat=g["p_feat"][off:off + n],
            pos=g["pos_init"][off:off + n], edge_index=g["bond_index"][:, sel] - off, edge_type=g["bond_type"][sel],
            smiles="rxn%d" % k)))
        off += n
    return g, out


def test_from_data_list_matches_synthetic_batch():
    g, data = _reactions([10, 17, 25, 12])
    b = Batch.from_data_list(data)
    assert b.num_graphs == 4 and b.num_nodes == 64
    for k_batch, k_syn in (("atom_type", "atom_type"), ("r_feat", "r_feat"), ("p_feat", "p_feat"), ("batch", "batch"),
                           ("edge_index", "bond_index"), ("edge_type", "bond_type"), ("pos", "pos_init")):
        assert torch.equal(b[k_batch], g[k_syn]), k_batch
    assert b.smiles == ["rxn0", "rxn1", "rxn2", "rxn3"]
    assert torch.equal(b.num_nodes_per_graph, torch.tensor([10, 17, 25, 12]))


def test_to_data_list_round_trip():
    _, data = _reactions([11, 23, 14])
    back = Batch.from_data_list(data).to_data_list()
    assert len(back) == 3
    for a, b in zip(data, back):
        assert a.num_nodes == b.num_nodes and a.smiles == b.smiles
        for k in ("atom_type", "r_feat", "p_feat", "pos", "edge_index", "edge_type"):
            assert torch.equal(a[k], b[k]), k


def test_batching_and_repeat_like_sampling_py():
    _, data = _reactions([10, 12, 14])
    chunks = list(batching(data, 4, repeat_num=3))
    assert [len(c) for c in chunks] == [4, 4, 1]
    assert [d.smiles for c in chunks for d in c] == ["rxn0"] * 3 + ["rxn1"] * 3 + ["rxn2"] * 3
    chunks[0][0].pos.zero_()  # repeat() clones: the source list is untouched
    assert float(data[0].pos.abs().sum()) > 0


def test_pickle_round_trip(tmp_path):
    _, data = _reactions([10, 12])
    for d in data:
        d.pos_gen = torch.randn(d.num_nodes, 3)
    p = os.path.join(tmp_path, "samples.pkl")
    save_samples(data, p)
    back = load_samples(p)
    assert len(back) == 2 and torch.equal(back[1].pos_gen, data[1].pos_gen) and back[0].smiles == "rxn0"


def test_rxn0_fixture_collates(rxn0):
    """The featurised rxn_0 graph (from the reference's own result pickle) through the shim."""
    d = count_nodes_per_graph(Data(atom_type=rxn0["atom_type"], r_feat=rxn0["r_feat"], p_feat=rxn0["p_feat"],
                                   edge_index=rxn0["bond_index"], edge_type=rxn0["bond_type"],
                                   pos=torch.zeros(13, 3), smiles="rxn_0"))
    b = Batch.from_data_list([d, d.clone()])
    assert b.num_nodes == 26 and b.edge_index.size(1) == 52
    assert torch.equal(b.edge_index[:, 26:], rxn0["bond_index"] + 13)
    assert torch.equal(b.batch, torch.tensor([0] * 13 + [1] * 13))


@pytest.mark.reference
def test_load_reference_result_pickle():
    path = "/root/reference/birkholz_benchmark/rxn_0/samples_all.pkl"
    if not os.path.isfile(path):
        pytest.skip("reference checkout not present")
    samples = load_samples(path)
    assert len(samples) == 100 and samples[0].atom_type.numel() == 13 and samples[0].pos_gen.shape == (13, 3)
    b = Batch.from_data_list(samples[:3])
    assert b.num_graphs == 3 and b.edge_index.size(1) == 78
    fx = torch.load(os.path.join(GOLDEN, "rxn0_graph.pt"), weights_only=False)
    assert torch.equal(samples[0].edge_index, fx["bond_index"])


class _StubSampler:
    """Duck-typed EnsembleSampler: records the dynamic_sampling call and returns a synthetic trajectory."""

    def __init__(self, n_timesteps=5000):
        self.alphas = torch.linspace(0.999, 0.01, n_timesteps)
        self.num_timesteps = n_timesteps
        self.calls = []

    def dynamic_sampling(self, **kw):
        self.calls.append(kw)
        n = kw["atom_type"].numel()
        traj = [torch.full((n, 3), float(k + 1)) for k in range(kw["n_steps"])] if kw.get("keep_traj", True) else []
        return torch.arange(n * 3, dtype=torch.float32).reshape(n, 3), traj


def test_sample_batch_host_logic_from_ts_guess_and_traj_scaling():
    """sampling.py:171-216 on the host side: the TS-guess start is divided by sqrt(alpha[start_t - 1]), the
    trajectory is scaled by sqrt(alpha) of the visited time indices (latest first), results split per reaction."""
    from tsdiff_b200.data import sample_batch
    _, data = _reactions([10, 12])
    batch = Batch.from_data_list(data)
    model = _StubSampler()
    res = sample_batch(model, batch, n_steps=4, sampling_type="ld", from_ts_guess=True, denoise_from_time_t=3000,
                       noise_from_time_t=1500, save_traj=True)
    call = model.calls[0]
    want_init = batch.pos / model.alphas[1499].sqrt()
    assert torch.allclose(call["pos_init"], want_init) and call["noise_from_time_t"] == 1500
    assert call["denoise_from_time_t"] == 3000 and call["extend_order"] is True and call["num_graphs"] == 2
    scale = model.alphas[3000 - 4:3000].flip(0).sqrt()
    assert len(res) == 2 and res[0].pos_gen.shape == (4, 10, 3) and res[1].pos_gen.shape == (4, 12, 3)
    for k in range(4):
        assert torch.allclose(res[0].pos_gen[k], torch.full((10, 3), float(k + 1)) * scale[k])
    # without save_traj: final positions split by reaction, default (random normal) start of the right shape
    res = sample_batch(model, batch, n_steps=3, sampling_type="ld")
    assert model.calls[1]["pos_init"].shape == (22, 3) and model.calls[1]["keep_traj"] is False
    assert torch.equal(res[1].pos_gen, torch.arange(66, dtype=torch.float32).reshape(22, 3)[10:])
    assert [r.smiles for r in res] == ["rxn0", "rxn1"]
