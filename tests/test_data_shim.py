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
