"""CPU: the oracle restatement against the golden vectors produced by the reference's own
Python (tests/golden/make_golden.py), plus property tests of the graph builder."""
import json
import os

import pytest
import torch

from oracle import tsdiff_oracle as O
from oracle import third_party as tp
from tsdiff_b200.config import QM9_DEFAULT_MODEL, TRAIN_CONFIG_MODEL
from tsdiff_b200.synthetic import make_batch

from conftest import DDPM_CASES, DUALENC_BRANCH_CASES, GOLDEN, graph_for
from helpers import make_model, max_rel_err, oracle_params, rel_err

FP32_TOL = 2e-5  # oracle vs reference on CPU: same op graph, only summation-order noise


def test_schedule_constants():
    betas, alphas = O.schedule_tensors(TRAIN_CONFIG_MODEL)
    sig = O.sigmas_of(alphas)
    ref = json.load(open(os.path.join(GOLDEN, "schedule.json")))
    assert abs(float(sig[0]) - ref["sigma_0"]) <= 1e-9
    assert abs(float(sig[-1]) - ref["sigma_last"]) <= 1e-6
    assert abs(float(sig[2500]) - ref["sigma_2500"]) <= 1e-7
    assert abs(float(betas[0]) - ref["beta_0"]) <= 1e-12 and abs(float(betas[-1]) - ref["beta_last"]) <= 1e-9
    # SURVEY.md 8a-10 constants
    assert abs(float(sig[0]) - 2.2509e-3) < 1e-6 and abs(float(sig[-1]) - 12.1685) < 1e-3


@pytest.mark.parametrize("seed,name", [(0, "condensenc_train_config_seed0"), (1, "condensenc_train_config_seed1")])
def test_state_dict_matches_reference_condensenc(seed, name):
    man = json.load(open(os.path.join(GOLDEN, "state_dict_manifests.json")))[name]
    sd = make_model("condensenc", seed).state_dict()
    assert list(sd.keys()) == list(man.keys()) and len(sd) == 158
    for k, v in sd.items():
        assert list(v.shape) == man[k]["shape"], k
        assert abs(float(v.double().sum()) - man[k]["sum"]) <= 1e-9 * max(1.0, abs(man[k]["sum"])), k
        assert abs(float(v.double().abs().sum()) - man[k]["abs_sum"]) <= 1e-9 * max(1.0, man[k]["abs_sum"]), k


def test_state_dict_matches_reference_dualenc():
    man = json.load(open(os.path.join(GOLDEN, "state_dict_manifests.json")))["dualenc_qm9_default_seed0"]
    sd = make_model("dualenc", 0).state_dict()
    assert list(sd.keys()) == list(man.keys()) and len(sd) == 198
    for k, v in sd.items():
        assert list(v.shape) == man[k]["shape"], k
        assert abs(float(v.double().sum()) - man[k]["sum"]) <= 1e-9 * max(1.0, abs(man[k]["sum"])), k


def test_aliased_parameters_share_storage():
    m = make_model("condensenc", 0)
    sd = m.state_dict()
    assert sd["model.0.bond_emb.weight"].data_ptr() == sd["edge_encoder.bond_emb.weight"].data_ptr()
    assert sum(p.numel() for p in m.parameters() if p.requires_grad) == 2770305
    a = make_model("dualenc", 0)
    assert sum(p.numel() for p in a.parameters() if p.requires_grad) == 793858


@pytest.mark.parametrize("case", ["graph_rxn0_o3", "graph_rxn0_o4", "graph_syn4_o3", "graph_syn4_o4"])
def test_graph_builder_bit_exact(case, golden, rxn0, syn4):
    g, ref = graph_for(case, rxn0, syn4), golden[case]
    order = int(case[-1])
    n = g["atom_type"].numel()
    loc, tr, tp_ = O.ts_bond_order_edges(n, g["bond_index"], g["bond_type"], order)
    assert torch.equal(loc, ref["local_index"]) and torch.equal(tr, ref["local_type_r"])
    assert torch.equal(tp_, ref["local_type_p"])
    glob, _ = O.union_with_radius_graph(ref["pos"], loc, tr, 10.0, g["batch"])
    assert torch.equal(glob, ref["global_index"])
    ia, ta = O.order_radius_graph(n, ref["pos"], g["bond_index"], g["bond_type"], g["batch"], order, 10.0)
    assert torch.equal(ia, ref["a_index"]) and torch.equal(ta, ref["a_type"])


def test_rxn0_edge_counts(rxn0):
    """order-4 local 132, order-3 local 106, all pairs 156; path A order-3: 116."""
    n = 13
    assert O.ts_bond_order_edges(n, rxn0["bond_index"], rxn0["bond_type"], 4)[0].size(1) == 132
    assert O.ts_bond_order_edges(n, rxn0["bond_index"], rxn0["bond_type"], 3)[0].size(1) == 106
    assert O.bond_order_edges(n, rxn0["bond_index"], rxn0["bond_type"], 3)[0].size(1) == 116


@pytest.mark.parametrize("seed", [0, 1, 2])
def test_hop_bfs_equals_dense_powers(seed):
    g = make_batch(3, seed=seed)
    n = g["atom_type"].numel()
    for order in (2, 3, 4):
        adj = tp.to_dense_adj(g["bond_index"], max_num_nodes=n).squeeze(0)
        dense = O._hop_order_dense(adj, order).long()
        assert torch.equal(dense, O.hop_order_bfs(n, g["bond_index"], order))


@pytest.mark.parametrize("case", ["b_rxn0_fwd", "b_rxn0_fwd_wide", "b_syn4_fwd", "b_syn4_fwd_wide"])
def test_condensenc_forward_matches_reference(case, golden, rxn0, syn4):
    g, ref = graph_for(case, rxn0, syn4), golden[case]
    p = oracle_params(make_model("condensenc", 0))
    ei, idx, ln = O.condensenc_forward(p, TRAIN_CONFIG_MODEL, g["atom_type"], g["r_feat"], g["p_feat"], ref["pos"],
                                       g["bond_index"], g["bond_type"], g["batch"])
    assert torch.equal(idx, ref["edge_index"])
    assert rel_err(ln, ref["edge_length"]) < 1e-6
    assert rel_err(ei, ref["edge_inv"]) < FP32_TOL and max_rel_err(ei, ref["edge_inv"]) < 1e-4


def test_known_answer_rxn0(golden):
    """SURVEY.md 8c known answer from the verbatim reference run."""
    ei = golden["b_rxn0_fwd"]["edge_inv"]
    assert ei.numel() == 156
    assert abs(float(ei.mean()) - (-0.04846735)) < 1e-7 and abs(float(ei.std()) - 0.01158494) < 1e-7


def test_ensemble_forward_matches_reference(golden, rxn0):
    ref = golden["b_rxn0_ens2_fwd"]
    ps = [oracle_params(make_model("condensenc", s)) for s in (0, 1)]
    ei, idx, _ = O.ensemble_forward(ps, TRAIN_CONFIG_MODEL, rxn0["atom_type"], rxn0["r_feat"], rxn0["p_feat"],
                                    ref["pos"], rxn0["bond_index"], rxn0["bond_type"], rxn0["batch"])
    assert torch.equal(idx, ref["edge_index"]) and rel_err(ei, ref["edge_inv"]) < FP32_TOL


@pytest.mark.parametrize("case,seeds", [("b_rxn0_ld20", (0,)), ("b_syn4_ld10", (0,)), ("b_rxn0_ens2_ld5", (0, 1))])
def test_ld_trajectory_matches_reference(case, seeds, golden, rxn0, syn4):
    g, ref = graph_for(case, rxn0, syn4), golden[case]
    ps = [oracle_params(make_model("condensenc", s)) for s in seeds]
    n_steps = ref["noise"].size(0)
    pos, traj = O.dynamic_sampling_ld(ps, TRAIN_CONFIG_MODEL, g["atom_type"], g["r_feat"], g["p_feat"],
                                      ref["pos_init"], g["bond_index"], g["bond_type"], g["batch"], n_steps, 1e-7,
                                      clip=1000, noise=ref["noise"])
    assert (torch.stack(traj) - ref["traj"]).abs().max() < 1e-4  # Angstrom, positions are O(10)
    assert (pos - ref["pos"]).abs().max() < 1e-4


@pytest.mark.parametrize("case", sorted(DDPM_CASES))
def test_sampler_branches_match_reference(case, golden_ddpm, rxn0, syn4):
    """sampler.py:149-182 starts (from_ts_guess noising, zero-noise, default) and the ddpm update (:215-236)."""
    g, ref, kw = graph_for(case, rxn0, syn4), golden_ddpm[case], DDPM_CASES[case]
    ps = [oracle_params(make_model("condensenc", 0))]
    pos, traj = O.dynamic_sampling(ps, TRAIN_CONFIG_MODEL, g["atom_type"], g["r_feat"], g["p_feat"], ref["pos_init"],
                                   g["bond_index"], g["bond_type"], g["batch"], ref["noise"].size(0), 1e-7, clip=1000,
                                   noise=ref["noise"], init_noise=ref.get("init_noise"), **kw)
    assert (torch.stack(traj) - ref["traj"]).abs().max() < 1e-4  # Angstrom
    assert (pos - ref["pos"]).abs().max() < 1e-4


@pytest.mark.parametrize("case", ["a_rxn0_fwd", "a_rxn0_fwd_wide", "a_syn4_fwd", "a_syn4_fwd_wide"])
def test_dualenc_forward_matches_reference(case, golden, rxn0, syn4):
    g, ref = graph_for(case, rxn0, syn4), golden[case]
    p = oracle_params(make_model("dualenc", 0))
    ig, il, idx, typ, ln, local = O.dualenc_forward(p, QM9_DEFAULT_MODEL, g["atom_type"], ref["pos"], g["bond_index"],
                                                     g["bond_type"], g["batch"])
    assert torch.equal(idx, ref["edge_index"]) and torch.equal(typ, ref["edge_type"])
    assert torch.equal(local, ref["local_edge_mask"])
    assert rel_err(ig, ref["edge_inv_global"]) < FP32_TOL and rel_err(il, ref["edge_inv_local"]) < FP32_TOL


def test_dualenc_embedding_renorm_side_effect(golden, rxn0):
    man = json.load(open(os.path.join(GOLDEN, "state_dict_manifests.json")))
    before = man["dualenc_qm9_default_seed0"]["encoder_global.node_emb.weight"]
    after = man["dualenc_qm9_default_seed0_after_fwd"]["encoder_global.node_emb.weight"]
    assert before["abs_sum"] != after["abs_sum"]
    p = oracle_params(make_model("dualenc", 0))
    O.dualenc_forward(p, QM9_DEFAULT_MODEL, rxn0["atom_type"], golden["a_rxn0_fwd"]["pos"], rxn0["bond_index"],
                      rxn0["bond_type"], rxn0["batch"])
    got = float(p["encoder_global.node_emb.weight"].double().abs().sum())
    assert abs(got - after["abs_sum"]) < 1e-4


@pytest.mark.parametrize("case", ["a_rxn0_ld10", "a_syn4_ld5"])
def test_dualenc_ld_matches_reference(case, golden, rxn0, syn4):
    g, ref = graph_for(case, rxn0, syn4), golden[case]
    p = oracle_params(make_model("dualenc", 0))
    kw = {k: float(ref[k]) for k in ("clip", "clip_local", "w_global") if k in ref}
    n_steps = ref["noise"].size(0)
    pos, traj = O.dualenc_ld_sample(p, QM9_DEFAULT_MODEL, g["atom_type"], ref["pos_init"], g["bond_index"],
                                    g["bond_type"], g["batch"], n_steps, 1e-7, noise=ref["noise"], **kw)
    assert (torch.stack(traj) - ref["traj"]).abs().max() < 2e-3
    assert (pos - ref["pos"]).abs().max() < 2e-3


def test_radius_cap_rule():
    """torch_cluster CUDA rule: per centre the first 33 in-range atoms by index (self
    included) survive; the graph becomes asymmetric when the cap binds."""
    torch.manual_seed(0)
    pos = torch.randn(50, 3) * 0.5
    idx = tp.radius_graph(pos, r=100.0, batch=torch.zeros(50, dtype=torch.long))
    deg_in = torch.bincount(idx[1], minlength=50)
    assert (deg_in[:33] == 32).all()  # 33 kept incl. self, self dropped (self is among the first 33)
    assert (deg_in[33:] == 33).all()  # centre >= 33: the first 33 atoms are 0..32, self not among them
    assert (idx[0][idx[1] == 40] == torch.arange(33)).all()


@pytest.mark.parametrize("case", sorted(DUALENC_BRANCH_CASES))
def test_dualenc_sampler_branches_match_reference(case, golden_dualenc_branches, rxn0, syn4):
    """dualenc.py:861-944: ddpm_noisy (the class default), ddpm_det, generalized."""
    g, ref, kw = graph_for(case, rxn0, syn4), golden_dualenc_branches[case], DUALENC_BRANCH_CASES[case]
    p = oracle_params(make_model("dualenc", 0))
    pos, traj = O.dualenc_ld_sample(p, QM9_DEFAULT_MODEL, g["atom_type"], ref["pos_init"], g["bond_index"],
                                    g["bond_type"], g["batch"], ref["noise"].size(0), 1e-7, noise=ref["noise"], **kw)
    scale = max(1.0, float(ref["traj"].abs().max()))  # the ddpm branches blow positions up at random init
    assert (torch.stack(traj) - ref["traj"]).abs().max() < 1e-4 * scale
    assert (pos - ref["pos"]).abs().max() < 1e-4 * scale


@pytest.mark.parametrize("name", ["syn4", "rxn0"])
def test_losses_match_reference(name, golden_loss, rxn0, syn4):
    """get_loss of both networks (condensenc.py:267-328, dualenc.py:425-562) with the reference's own draws."""
    g = graph_for(name, rxn0, syn4)
    ref = golden_loss["b_" + name]
    loss = O.condensenc_loss(oracle_params(make_model("condensenc", 0)), TRAIN_CONFIG_MODEL, g["atom_type"], g["r_feat"],
                             g["p_feat"], ref["pos"], g["bond_index"], g["bond_type"], g["batch"], ref["time_step"],
                             ref["pos_noise"])
    assert loss.shape == ref["loss"].shape and rel_err(loss, ref["loss"]) < 1e-4
    ref = golden_loss["a_" + name]
    loss, lg, ll = O.dualenc_loss(oracle_params(make_model("dualenc", 0)), QM9_DEFAULT_MODEL, g["atom_type"], ref["pos"],
                                  g["bond_index"], g["bond_type"], g["batch"], ref["time_step"], ref["pos_noise"])
    assert rel_err(loss, ref["loss"]) < 1e-4 and rel_err(lg, ref["loss_global"]) < 1e-4
    assert rel_err(ll, ref["loss_local"]) < 1e-4


def test_oracle_autograd_matches_reference_gradients(golden_loss, syn4):
    """The oracle is functional PyTorch, so its autograd is the checker for the (not yet built) backward
    kernels: gradients of get_loss(...).mean() (train.py:128-142) against the reference's own autograd."""
    gold = json.load(open(os.path.join(GOLDEN, "golden_grads.json")))
    ref = golden_loss["b_syn4"]
    m = make_model("condensenc", 0)
    p = {k: v.clone().requires_grad_(v.is_floating_point() and k not in ("betas", "alphas"))
         for k, v in oracle_params(m).items()}
    loss = O.condensenc_loss(p, TRAIN_CONFIG_MODEL, syn4["atom_type"], syn4["r_feat"], syn4["p_feat"], ref["pos"],
                             syn4["bond_index"], syn4["bond_type"], syn4["batch"], ref["time_step"],
                             ref["pos_noise"]).mean()
    assert abs(float(loss.detach()) - gold["loss_mean"]) < 1e-4 * gold["loss_mean"]
    loss.backward()
    checked = 0
    for name, want in gold["params"].items():
        g = p[name].grad
        assert g is not None, name
        assert abs(float(g.double().norm()) - want["norm"]) <= 2e-4 * max(want["norm"], 1e-6), name
        checked += 1
    assert checked == 80
    for name, flat in gold["full"].items():
        want = torch.tensor(flat).reshape(p[name].shape)
        assert rel_err(p[name].grad, want) < 2e-4, name
