"""GPU parity tests proper: every call goes through the C-ABI (ctypes) and is compared with
the CPU oracle on the same seeded inputs and with the golden vectors produced by the
reference's own Python.  Bit-exact for index / integer work; fp32 tolerance stated inline."""
import ctypes as C

import numpy as np
import pytest
import torch

from oracle import philox as OP
from oracle import third_party as tp
from oracle import tsdiff_oracle as O
from tsdiff_b200 import _lib as L
from tsdiff_b200 import engine as E
from tsdiff_b200.config import QM9_DEFAULT_MODEL, TRAIN_CONFIG_MODEL
from tsdiff_b200.synthetic import make_batch, shard_batch

from conftest import DDPM_CASES, DUALENC_BRANCH_CASES, graph_for
from helpers import make_model, max_rel_err, oracle_params, rel_err, to_dev

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
EPS_TOL = 1e-4  # north_star: per-step eps within 1e-4 relative in fp32


def _plan_edges(plan):
    e = plan.edge_count()
    return e, torch.stack([plan.row[:e], plan.col[:e]]).long().cpu()


def _check_csr(plan):
    e = plan.edge_count()
    row, col = plan.row[:e].long().cpu(), plan.col[:e].long().cpu()
    n = plan.num_nodes
    row_ptr, in_ptr, in_eid = plan.row_ptr.long().cpu(), plan.in_ptr.long().cpu(), plan.in_eid[:e].long().cpu()
    assert int(row_ptr[n]) == e and int(in_ptr[n]) == e
    assert torch.equal(torch.bincount(row, minlength=n), row_ptr[1:] - row_ptr[:-1])
    assert torch.equal(torch.bincount(col, minlength=n), in_ptr[1:] - in_ptr[:-1])
    key = row * n + col
    assert bool((key[1:] > key[:-1]).all()), "edge list not strictly row-major sorted"
    # in-CSR: a permutation of the edges, grouped by col, sources ascending
    assert torch.equal(torch.sort(in_eid)[0], torch.arange(e))
    kin = col[in_eid] * n + row[in_eid]
    assert bool((kin[1:] > kin[:-1]).all())


@pytest.mark.parametrize("name,scale", [("rxn0", 1.0), ("rxn0", 5.0), ("syn4", 3.0), ("syn4", 6.0), ("syn4", 12.0)])
def test_edge_build_path_b_bit_exact(name, scale, rxn0, syn4):
    g = graph_for(name, rxn0, syn4)
    torch.manual_seed(3)
    pos = torch.randn(g["atom_type"].numel(), 3) * scale
    d = to_dev(g, DEV)
    plan = E.BatchPlan(0, d["batch"], d["bond_index"], d["bond_type"], 4, 3)
    plan.build_edges(pos.to(DEV).contiguous(), 10.0)
    e, idx = _plan_edges(plan)
    ref_idx, ref_r, ref_p = O.condensed_graph(pos, g["bond_index"], g["bond_type"], g["batch"], 4, 10.0)
    assert torch.equal(idx, ref_idx)
    code = plan.tab0[:e].cpu()
    assert torch.equal((code & 0xFFFF).long(), ref_r) and torch.equal((code >> 16).long(), ref_p)
    assert torch.equal(plan.length[:e].cpu(), O.edge_lengths(pos, ref_idx))  # bit-exact fp32
    # second graph (pred_edge_order = 3) as a mask over the first
    ref3, r3, p3 = O.condensed_graph(pos, g["bond_index"], g["bond_type"], g["batch"], 3, 10.0)
    sel = plan.in_b[:e].bool().cpu()
    assert torch.equal(idx[:, sel], ref3)
    code3 = plan.tab1[:e].cpu()[sel]
    assert torch.equal((code3 & 0xFFFF).long(), r3) and torch.equal((code3 >> 16).long(), p3)
    _check_csr(plan)


@pytest.mark.parametrize("name,scale", [("rxn0", 1.0), ("rxn0", 5.0), ("syn4", 6.0)])
def test_edge_build_path_a_bit_exact(name, scale, rxn0, syn4):
    g = graph_for(name, rxn0, syn4)
    torch.manual_seed(4)
    n = g["atom_type"].numel()
    pos = torch.randn(n, 3) * scale
    d = to_dev(g, DEV)
    plan = E.BatchPlan(1, d["batch"], d["bond_index"], d["bond_type"], 3, 0)
    plan.build_edges(pos.to(DEV).contiguous(), 10.0)
    e, idx = _plan_edges(plan)
    ref_idx, ref_t = O.order_radius_graph(n, pos, g["bond_index"], g["bond_type"], g["batch"], 3, 10.0)
    assert torch.equal(idx, ref_idx) and torch.equal(plan.tab0[:e].long().cpu(), ref_t)
    rows, _ = O.dualenc_type_decode(ref_t, False)
    assert torch.equal((plan.tab1[:e].cpu() & 0xFFFF).long(), rows)
    _check_csr(plan)


def test_edge_build_neighbor_cap_asymmetric():
    """Stress shape: ~60-atom graphs, enlarged cutoff -> the 32-neighbour cap binds and the
    radius graph is asymmetric."""
    g = make_batch(5, seed=9, min_atoms=55, max_atoms=65)
    torch.manual_seed(5)
    pos = torch.randn(g["atom_type"].numel(), 3) * 4.0
    d = to_dev(g, DEV)
    plan = E.BatchPlan(0, d["batch"], d["bond_index"], d["bond_type"], 4, 3)
    plan.build_edges(pos.to(DEV).contiguous(), 15.0)
    e, idx = _plan_edges(plan)
    ref_idx, ref_r, _ = O.condensed_graph(pos, g["bond_index"], g["bond_type"], g["batch"], 4, 15.0)
    assert torch.equal(idx, ref_idx)
    n = plan.num_nodes
    key = set((idx[0] * n + idx[1]).tolist())
    assert any((int(c) * n + int(r)) not in key for r, c in idx.t().tolist()), "expected an asymmetric edge set"
    _check_csr(plan)


def test_edge_build_large_batch_properties():
    """BASELINE config-2 size (batch 100): size-independent properties + oracle equality."""
    g = make_batch(100, seed=0)
    pos = g["pos_init"] * 4.0
    d = to_dev(g, DEV)
    plan = E.BatchPlan(0, d["batch"], d["bond_index"], d["bond_type"], 4, 3)
    for _ in range(2):  # idempotent
        plan.build_edges(pos.to(DEV).contiguous(), 10.0)
    e, idx = _plan_edges(plan)
    _check_csr(plan)
    assert bool((g["batch"][idx[0]] == g["batch"][idx[1]]).all()), "edge across two reactions"
    ref_idx, _, _ = O.condensed_graph(pos, g["bond_index"], g["bond_type"], g["batch"], 4, 10.0)
    assert torch.equal(idx, ref_idx)


def test_bond_table_rejects_cross_graph_bond(syn4):
    d = to_dev(syn4, DEV)
    bad = d["bond_index"].clone()
    bad[1, 0] = d["atom_type"].numel() - 1
    with pytest.raises(ValueError):
        E.BatchPlan(0, d["batch"], bad, d["bond_type"], 4, 3)


def test_philox_normals_match_numpy():
    lib = L.load()
    out = torch.empty(1000, 3, device=DEV)
    L.check(lib.tsd_philox_normal(1000, 0x1234567890ABCDEF, 17, 5_000_000_000, L.ptr(out),
                                  C.c_void_p(torch.cuda.current_stream().cuda_stream)), "philox")
    ref = OP.normals(1000, 0x1234567890ABCDEF, 17, 5_000_000_000)
    assert np.abs(out.cpu().numpy() - ref).max() < 2e-5
    big = torch.empty(200000, 3, device=DEV)
    L.check(lib.tsd_philox_normal(200000, 7, 0, 0, L.ptr(big), C.c_void_p(torch.cuda.current_stream().cuda_stream)),
            "philox")
    assert abs(float(big.mean())) < 5e-3 and abs(float(big.std()) - 1.0) < 5e-3


@pytest.mark.parametrize("case", ["b_rxn0_fwd", "b_rxn0_fwd_wide", "b_syn4_fwd", "b_syn4_fwd_wide"])
def test_condensenc_forward_vs_reference_golden(case, golden, rxn0, syn4):
    g, ref = graph_for(case, rxn0, syn4), golden[case]
    m = make_model("condensenc", 0, DEV)
    d = to_dev(g, DEV)
    ei, idx, ln = m(d["atom_type"], d["r_feat"], d["p_feat"], ref["pos"].to(DEV), d["bond_index"], d["bond_type"],
                    d["batch"], torch.zeros(g["num_graphs"], dtype=torch.long, device=DEV))
    assert ei.shape == ref["edge_inv"].shape and idx.dtype == torch.long
    assert torch.equal(idx.cpu(), ref["edge_index"])
    assert rel_err(ln, ref["edge_length"]) < 1e-6
    assert rel_err(ei, ref["edge_inv"]) < EPS_TOL and max_rel_err(ei, ref["edge_inv"]) < EPS_TOL


def test_condensenc_forward_vs_oracle_batch100():
    g = make_batch(100, seed=0)
    m = make_model("condensenc", 0, DEV)
    d = to_dev(g, DEV)
    pos = g["pos_init"] * 5.0
    ei, idx, ln = m(d["atom_type"], d["r_feat"], d["p_feat"], pos.to(DEV), d["bond_index"], d["bond_type"],
                    d["batch"], None)
    rei, ridx, rln = O.condensenc_forward(oracle_params(m), TRAIN_CONFIG_MODEL, g["atom_type"], g["r_feat"],
                                          g["p_feat"], pos, g["bond_index"], g["bond_type"], g["batch"])
    assert torch.equal(idx.cpu(), ridx) and torch.equal(ln.cpu(), rln)
    assert rel_err(ei, rei) < EPS_TOL and max_rel_err(ei, rei) < EPS_TOL


def test_ensemble_forward_vs_reference_golden(golden, rxn0):
    from tsdiff_b200.models.sampler import EnsembleSampler
    ref = golden["b_rxn0_ens2_fwd"]
    ens = EnsembleSampler([make_model("condensenc", s, DEV) for s in (0, 1)])
    d = to_dev(rxn0, DEV)
    ei, idx, ln = ens(d["atom_type"], d["r_feat"], d["p_feat"], ref["pos"].to(DEV), d["bond_index"], d["bond_type"],
                      d["batch"], None)
    assert torch.equal(idx.cpu(), ref["edge_index"])
    assert rel_err(ei, ref["edge_inv"]) < EPS_TOL and max_rel_err(ei, ref["edge_inv"]) < EPS_TOL


def test_eq_transform_kernel_vs_oracle(syn4):
    lib = L.load()
    torch.manual_seed(8)
    n = syn4["atom_type"].numel()
    pos = torch.randn(n, 3) * 4.0
    d = to_dev(syn4, DEV)
    plan = E.BatchPlan(0, d["batch"], d["bond_index"], d["bond_type"], 4, 3)
    posd = pos.to(DEV).contiguous()
    plan.build_edges(posd, 10.0)
    e, idx = _plan_edges(plan)
    score = torch.randn(e)
    sd = torch.zeros(plan.edge_capacity, device=DEV)
    sd[:e] = score.to(DEV)
    ch = L.ScoreChannel(sd.data_ptr(), None, 0, 0.0, 1.0)
    out = torch.empty(n, 3, device=DEV)
    L.check(lib.tsd_eq_transform(C.byref(plan.c_batch), C.byref(plan.c_edges), L.ptr(posd), C.byref(ch), 1.0,
                                 L.ptr(out), C.c_void_p(torch.cuda.current_stream().cuda_stream)), "eq_transform")
    ref = O.eq_transform(score.unsqueeze(-1), pos, idx, O.edge_lengths(pos, idx).unsqueeze(-1))
    assert torch.equal(out.cpu(), ref), "eq_transform must be bit-exact (same association order)"


@pytest.mark.parametrize("case,seeds", [("b_rxn0_ld20", (0,)), ("b_syn4_ld10", (0,)), ("b_rxn0_ens2_ld5", (0, 1))])
@pytest.mark.parametrize("use_graph", [False, True])
def test_dynamic_sampling_vs_reference_golden(case, seeds, use_graph, golden, rxn0, syn4):
    from tsdiff_b200.models.sampler import EnsembleSampler
    g, ref = graph_for(case, rxn0, syn4), golden[case]
    ens = EnsembleSampler([make_model("condensenc", s, DEV) for s in seeds])
    d = to_dev(g, DEV)
    n_steps = ref["noise"].size(0)
    pos, traj = ens.dynamic_sampling(d["atom_type"], d["r_feat"], d["p_feat"], ref["pos_init"].to(DEV),
                                     d["bond_index"], d["bond_type"], d["batch"], g["num_graphs"], extend_order=True,
                                     n_steps=n_steps, step_lr=1e-7, clip=1000, sampling_type="ld",
                                     noise=ref["noise"], use_graph=use_graph)
    assert len(traj) == n_steps and not traj[0].is_cuda and pos.is_cuda
    # stated bound (fp32 path): 1e-4 Angstrom over <= 100 steps
    assert (torch.stack(traj) - ref["traj"]).abs().max() < 1e-4
    assert (pos.cpu() - ref["pos"]).abs().max() < 1e-4


@pytest.mark.parametrize("case", sorted(DDPM_CASES))
@pytest.mark.parametrize("use_graph", [False, True])
def test_sampler_branches_vs_reference_golden(case, use_graph, golden_ddpm, rxn0, syn4):
    """The `ddpm` update (sampler.py:215-236, the reference's default sampling_type), the
    from_ts_guess noising / zero-noise starts (:149-177) and clip_pos, against trajectories
    produced by the reference's own Python with the same injected noise."""
    from tsdiff_b200.models.sampler import EnsembleSampler
    g, ref, kw = graph_for(case, rxn0, syn4), golden_ddpm[case], DDPM_CASES[case]
    ens = EnsembleSampler([make_model("condensenc", 0, DEV)])
    d = to_dev(g, DEV)
    n_steps = ref["noise"].size(0)
    pos, traj = ens.dynamic_sampling(d["atom_type"], d["r_feat"], d["p_feat"], ref["pos_init"].to(DEV),
                                     d["bond_index"], d["bond_type"], d["batch"], g["num_graphs"], extend_order=True,
                                     n_steps=n_steps, step_lr=1e-7, clip=1000, noise=ref["noise"],
                                     init_noise=ref.get("init_noise"), use_graph=use_graph, **kw)
    assert len(traj) == n_steps
    assert (torch.stack(traj) - ref["traj"]).abs().max() < 1e-4  # Angstrom (fp32 path)
    assert (pos.cpu() - ref["pos"]).abs().max() < 1e-4


@pytest.mark.parametrize("save_traj", [False, True])
def test_sampling_script_body_vs_oracle(save_traj, syn4):
    """tsdiff_b200.data.sample_batch = the loop body of sampling.py:169-225 (collate, sample, alpha-scaled
    trajectory, per-reaction results) against the oracle sampler on the same noise."""
    from tsdiff_b200.data import Batch, Data, count_nodes_per_graph, sample_batch
    from tsdiff_b200.models.sampler import EnsembleSampler
    sizes, off, data = [10, 17, 25, 12], 0, []
    for k, n in enumerate(sizes):
        sel = (syn4["bond_index"][0] >= off) & (syn4["bond_index"][0] < off + n)
        data.append(count_nodes_per_graph(Data(
            atom_type=syn4["atom_type"][off:off + n], r_feat=syn4["r_feat"][off:off + n],
            p_feat=syn4["p_feat"][off:off + n], pos=syn4["pos_init"][off:off + n],
            edge_index=syn4["bond_index"][:, sel] - off, edge_type=syn4["bond_type"][sel], smiles="rxn%d" % k)))
        off += n
    m = make_model("condensenc", 0, DEV)
    ens = EnsembleSampler([m])
    n_steps = 6
    gen = torch.Generator().manual_seed(77)
    noise = torch.randn(n_steps, 64, 3, generator=gen)
    pos_init = syn4["pos_init"]
    batch = Batch.from_data_list(data).to(DEV)
    results = sample_batch(ens, batch, n_steps=n_steps, step_lr=1e-7, clip=1000.0, sampling_type="ld",
                           save_traj=save_traj, pos_init=pos_init.to(DEV), noise=noise)
    ref_pos, ref_traj = O.dynamic_sampling(
        [oracle_params(m)], TRAIN_CONFIG_MODEL, syn4["atom_type"], syn4["r_feat"], syn4["p_feat"], pos_init,
        syn4["bond_index"], syn4["bond_type"], syn4["batch"], n_steps, 1e-7, clip=1000, noise=noise)
    assert len(results) == 4 and [r.smiles for r in results] == ["rxn0", "rxn1", "rxn2", "rxn3"]
    alphas = oracle_params(m)["alphas"]
    scale = alphas[alphas.numel() - n_steps:].flip(0).view(-1, 1, 1).sqrt()
    off = 0
    for r, n in zip(results, sizes):
        assert not r.pos_gen.is_cuda and r.atom_type.numel() == n
        if save_traj:
            want = (torch.stack(ref_traj) * scale)[:, off:off + n]
            assert r.pos_gen.shape == (n_steps, n, 3)
        else:
            want = ref_pos[off:off + n]
        assert (r.pos_gen - want).abs().max() < 1e-4
        off += n


def test_unknown_sampling_type_raises(rxn0):
    from tsdiff_b200.models.sampler import EnsembleSampler
    ens = EnsembleSampler([make_model("condensenc", 0, DEV)])
    d = to_dev(rxn0, DEV)
    with pytest.raises(NotImplementedError):
        ens.dynamic_sampling(d["atom_type"], d["r_feat"], d["p_feat"], torch.randn(13, 3, device=DEV), d["bond_index"],
                             d["bond_type"], d["batch"], 1, extend_order=True, n_steps=2, sampling_type="generalized")


def test_dynamic_sampling_nan_raises(rxn0):
    from tsdiff_b200.models.sampler import EnsembleSampler
    ens = EnsembleSampler([make_model("condensenc", 0, DEV)])
    d = to_dev(rxn0, DEV)
    pos_init = torch.randn(13, 3, device=DEV)
    noise = torch.zeros(3, 13, 3)
    noise[1, 4, 0] = float("nan")
    with pytest.raises(FloatingPointError):
        ens.dynamic_sampling(d["atom_type"], d["r_feat"], d["p_feat"], pos_init, d["bond_index"], d["bond_type"],
                             d["batch"], 1, extend_order=True, n_steps=3, step_lr=1e-7, sampling_type="ld",
                             noise=noise)


def test_philox_sampling_is_shard_invariant():
    """1-vs-2 shard equivalence: per-reaction results identical when the batch is split
    (no collective on the data path; noise keyed by global atom id)."""
    from tsdiff_b200.models.sampler import EnsembleSampler
    g = make_batch(6, seed=5)
    m = make_model("condensenc", 0, DEV)

    def run(data, offset):
        d = to_dev(data, DEV)
        ens = EnsembleSampler([m])
        pos, _ = ens.dynamic_sampling(d["atom_type"], d["r_feat"], d["p_feat"], d["pos_init"], d["bond_index"],
                                      d["bond_type"], d["batch"], data["num_graphs"], extend_order=True, n_steps=8,
                                      step_lr=1e-7, sampling_type="ld", seed=77, atom_offset=offset, keep_traj=False)
        return pos.cpu()

    full = run(g, 0)
    parts = []
    for r in range(2):
        sh = shard_batch(g, r, 2)
        parts.append(run(sh, sh["atom_offset"]))
    assert torch.equal(torch.cat(parts), full)


@pytest.mark.parametrize("case", ["a_rxn0_fwd", "a_rxn0_fwd_wide", "a_syn4_fwd", "a_syn4_fwd_wide"])
def test_dualenc_forward_vs_reference_golden(case, golden, rxn0, syn4):
    g, ref = graph_for(case, rxn0, syn4), golden[case]
    m = make_model("dualenc", 0, DEV)
    d = to_dev(g, DEV)
    out = m(d["atom_type"], ref["pos"].to(DEV), d["bond_index"], d["bond_type"], d["batch"], None, return_edges=True)
    ig, il, idx, typ, ln, local = out
    assert torch.equal(idx.cpu(), ref["edge_index"]) and torch.equal(typ.cpu(), ref["edge_type"])
    assert torch.equal(local.cpu(), ref["local_edge_mask"])
    assert rel_err(ig, ref["edge_inv_global"]) < EPS_TOL and rel_err(il, ref["edge_inv_local"]) < EPS_TOL
    assert max_rel_err(ig, ref["edge_inv_global"]) < EPS_TOL and max_rel_err(il, ref["edge_inv_local"]) < EPS_TOL


def test_dualenc_embedding_renorm_written_back(golden, rxn0):
    m = make_model("dualenc", 0, DEV)
    ref_p = oracle_params(m)
    d = to_dev(rxn0, DEV)
    m(d["atom_type"], golden["a_rxn0_fwd"]["pos"].to(DEV), d["bond_index"], d["bond_type"], d["batch"], None)
    O.embedding_renorm_(ref_p["encoder_global.node_emb.weight"], rxn0["atom_type"], 10.0)
    got = m.state_dict()["encoder_global.node_emb.weight"].cpu()
    assert rel_err(got, ref_p["encoder_global.node_emb.weight"]) < 1e-6
    assert float(got[6].norm()) <= 10.0 + 1e-4 and float(got[1].norm()) <= 10.0 + 1e-4


@pytest.mark.parametrize("case", ["a_rxn0_ld10", "a_syn4_ld5"])
def test_dualenc_langevin_vs_reference_golden(case, golden, rxn0, syn4):
    g, ref = graph_for(case, rxn0, syn4), golden[case]
    m = make_model("dualenc", 0, DEV)
    d = to_dev(g, DEV)
    kw = {k: float(ref[k]) for k in ("clip", "clip_local", "w_global") if k in ref}
    n_steps = ref["noise"].size(0)
    pos, traj = m.langevin_dynamics_sample(d["atom_type"], ref["pos_init"].to(DEV), d["bond_index"], d["bond_type"],
                                           d["batch"], g["num_graphs"], extend_order=True, n_steps=n_steps,
                                           step_lr=1e-7, sampling_type="ld", noise=ref["noise"], **kw)
    # random-init path A amplifies perturbations (|score| ~ 1e3): 2e-3 Angstrom stated
    assert (torch.stack(traj) - ref["traj"]).abs().max() < 2e-3
    assert (pos.cpu() - ref["pos"]).abs().max() < 2e-3


def test_library_is_the_compute_path():
    """The product must not route through the oracle or torch math: the parameter modules
    have no PyTorch forward, and no product source file mentions the oracle package."""
    import glob
    import os
    import tsdiff_b200
    import tsdiff_b200.models.layers as layers
    with pytest.raises(RuntimeError):
        layers.MultiLayerPerceptron(4, [4])(torch.zeros(1, 4))
    root = os.path.dirname(tsdiff_b200.__file__)
    for path in glob.glob(os.path.join(root, "**", "*.py"), recursive=True):
        text = open(path).read()
        assert "import oracle" not in text and "from oracle" not in text, path


@pytest.mark.parametrize("rows,k,n,act", [(1000, 256, 256, "ssp"), (28442, 256, 256, "swish"), (129, 512, 256, "none"),
                                          (5000, 128, 128, "relu"), (777, 128, 64, "none"), (20000, 256, 128, "none")])
def test_linear_kernel_vs_torch_fp32(rows, k, n, act):
    """The GEMM building block against a plain PyTorch fp32 reference of the same op
    (fp64-accumulated to make the reference the tighter side)."""
    lib = L.load()
    torch.manual_seed(rows)
    x = torch.randn(rows, k, device=DEV)
    w = torch.randn(n, k, device=DEV) / k ** 0.5
    b = torch.randn(n, device=DEV)
    out = torch.full((rows + 3, n), 7.0, device=DEV)  # 3 guard rows must stay untouched
    rows_dev = torch.tensor([rows], dtype=torch.int32, device=DEV)
    lin = L.linear(w, b)
    L.check(lib.tsd_linear(rows + 3, L.ptr(rows_dev), L.ptr(x), C.byref(lin), L.ACT[act], L.ptr(out), 0,
                           C.c_void_p(torch.cuda.current_stream().cuda_stream)), "tsd_linear")
    ref = (x.double() @ w.double().t() + b.double())
    ref = {"ssp": lambda t: torch.nn.functional.softplus(t) - np.log(2.0), "swish": lambda t: t * torch.sigmoid(t),
           "relu": torch.relu, "none": lambda t: t}[act](ref)
    assert rel_err(out[:rows], ref) < 2e-6
    assert bool((out[rows:] == 7.0).all())


def test_cfconv_aggregate_bit_exact_vs_sequential_scatter(syn4):
    lib = L.load()
    torch.manual_seed(9)
    n = syn4["atom_type"].numel()
    pos = torch.randn(n, 3) * 4.0
    d = to_dev(syn4, DEV)
    plan = E.BatchPlan(0, d["batch"], d["bond_index"], d["bond_type"], 4, 3)
    plan.build_edges(pos.to(DEV).contiguous(), 10.0)
    e, idx = _plan_edges(plan)
    h = 256
    x1 = torch.randn(n, h)
    filt = torch.randn(plan.edge_capacity, h)
    agg = torch.empty(n, h, device=DEV)
    L.check(lib.tsd_cfconv_aggregate(C.byref(plan.c_batch), C.byref(plan.c_edges), h, L.ptr(x1.to(DEV)),
                                     L.ptr(filt.to(DEV)), L.ptr(agg),
                                     C.c_void_p(torch.cuda.current_stream().cuda_stream)), "aggregate")
    ref = tp.scatter_add(x1[idx[0]] * filt[:e], idx[1], dim=0, dim_size=n)
    assert torch.equal(agg.cpu(), ref)


def test_cfconv_aggregate_staged_kernel_bit_exact_at_stress_size():
    """BASELINE config 5 shape: 80 reactions of ~60 atoms, cutoff 15 A -> edge capacity > 2^18, where
    tsd_cfconv_aggregate switches to the graph-staged kernel (x1 rows in shared memory, prefetched
    in-CSR).  Same sums in the same order as a sequential scatter_add: bit-exact."""
    lib = L.load()
    g = make_batch(80, seed=33, min_atoms=55, max_atoms=65)
    n = g["atom_type"].numel()
    torch.manual_seed(10)
    pos = torch.randn(n, 3) * 5.0
    d = to_dev(g, DEV)
    plan = E.BatchPlan(0, d["batch"], d["bond_index"], d["bond_type"], 4, 3)
    assert plan.edge_capacity >= (1 << 18)
    plan.build_edges(pos.to(DEV).contiguous(), 15.0)
    e, idx = _plan_edges(plan)
    h = 256
    x1 = torch.randn(n, h)
    filt = torch.randn(e, h)
    filt_dev = torch.zeros(plan.edge_capacity, h, device=DEV)
    filt_dev[:e] = filt.to(DEV)
    agg = torch.empty(n, h, device=DEV)
    L.check(lib.tsd_cfconv_aggregate(C.byref(plan.c_batch), C.byref(plan.c_edges), h, L.ptr(x1.to(DEV)),
                                     L.ptr(filt_dev), L.ptr(agg),
                                     C.c_void_p(torch.cuda.current_stream().cuda_stream)), "aggregate")
    ref = tp.scatter_add(x1[idx[0]] * filt, idx[1], dim=0, dim_size=n)
    assert torch.equal(agg.cpu(), ref)


def test_stress_shape_forward_vs_oracle():
    """BASELINE config 5 shape at reduced count: ~60-atom reactions, cutoff enlarged to 15 A so the
    32-neighbour cap binds (asymmetric radius graph).  Full path-B forward vs the oracle."""
    from tsdiff_b200.config import AttrDict
    cfg = AttrDict(dict(TRAIN_CONFIG_MODEL))
    cfg.edge_cutoff = 15.0
    cfg.encoder = AttrDict(dict(TRAIN_CONFIG_MODEL.encoder))
    cfg.encoder.cutoff = 15.0
    from tsdiff_b200.models.epsnet import get_model
    torch.manual_seed(0)
    m = get_model(cfg).to(DEV)
    g = make_batch(6, seed=21, min_atoms=55, max_atoms=65)
    d = to_dev(g, DEV)
    torch.manual_seed(6)
    pos = torch.randn(g["atom_type"].numel(), 3) * 4.0
    ei, idx, ln = m(d["atom_type"], d["r_feat"], d["p_feat"], pos.to(DEV), d["bond_index"], d["bond_type"], d["batch"], None)
    rei, ridx, rln = O.condensenc_forward(oracle_params(m), cfg, g["atom_type"], g["r_feat"], g["p_feat"], pos,
                                          g["bond_index"], g["bond_type"], g["batch"])
    assert torch.equal(idx.cpu(), ridx) and torch.equal(ln.cpu(), rln)
    n = g["atom_type"].numel()
    keys = set((ridx[0] * n + ridx[1]).tolist())
    assert any((int(c) * n + int(r)) not in keys for r, c in ridx.t().tolist()), "cap should make the graph asymmetric"
    assert rel_err(ei, rei) < EPS_TOL and max_rel_err(ei, rei) < EPS_TOL


@pytest.mark.parametrize("name,scale,cutoff", [("rxn0", 1.0, 10.0), ("syn4", 3.0, 10.0), ("syn4", 8.0, 10.0), ("stress", 4.0, 15.0)])
def test_undirected_pair_list_consistent(name, scale, cutoff, rxn0, syn4):
    """K2's unordered pair list: one entry per {i<j} with at least one directed edge, sorted, with the
    bit-identical length / table values of its directions (also when the neighbour cap makes the directed
    graph asymmetric), and edge_upair / in_upair mapping every directed edge / in-slot to its pair."""
    g = make_batch(5, seed=9, min_atoms=55, max_atoms=65) if name == "stress" else graph_for(name, rxn0, syn4)
    torch.manual_seed(12)
    pos = torch.randn(g["atom_type"].numel(), 3) * scale
    d = to_dev(g, DEV)
    plan = E.BatchPlan(0, d["batch"], d["bond_index"], d["bond_type"], 4, 3, upairs=True)
    assert plan.upairs
    plan.build_edges(pos.to(DEV).contiguous(), cutoff)
    e, idx = _plan_edges(plan)
    u = plan.work_count()
    n = plan.num_nodes
    urow, ucol = plan.u_row[:u].long().cpu(), plan.u_col[:u].long().cpu()
    lo, hi = torch.minimum(idx[0], idx[1]), torch.maximum(idx[0], idx[1])
    want = torch.unique(lo * n + hi)  # sorted
    assert torch.equal(urow * n + ucol, want)
    up = plan.edge_upair[:e].long().cpu()
    assert torch.equal(urow[up], lo) and torch.equal(ucol[up], hi)
    assert torch.equal(plan.u_length[:u].cpu()[up], plan.length[:e].cpu())          # bit-identical
    assert torch.equal(plan.u_tab0[:u].cpu()[up], plan.tab0[:e].cpu())
    assert torch.equal(plan.u_tab1[:u].cpu()[up], plan.tab1[:e].cpu())
    assert torch.equal(plan.in_upair[:e].cpu().long(), up[plan.in_eid[:e].long().cpu()])
    if name == "stress":
        assert u * 2 > e, "expected one-directional edges under the neighbour cap"


def test_asymmetric_bond_list_disables_pair_sharing(syn4):
    d = to_dev(syn4, DEV)
    keep = torch.ones(d["bond_index"].size(1), dtype=torch.bool, device=DEV)
    keep[0] = False  # drop one direction of a bond: tables are no longer symmetric
    plan = E.BatchPlan(0, d["batch"], d["bond_index"][:, keep], d["bond_type"][keep], 4, 3, upairs=True)
    assert not plan.upairs and plan.c_work_edges is plan.c_edges


def test_member_per_gpu_ensemble():
    """BASELINE config 3: one ensemble member per GPU, per-step all-reduce of the (N,3) scores; needs 2 GPUs."""
    import os
    import subprocess
    import sys
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs (gpurun --gpus 2)")
    script = os.path.join(os.path.dirname(os.path.abspath(__file__)), "mgpu_ensemble_check.py")
    port = 29600 + os.getpid() % 300
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                          "--master-addr", "127.0.0.1", "--master-port", str(port), script],
                         capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-2000:]


@pytest.mark.parametrize("case", sorted(DUALENC_BRANCH_CASES))
def test_dualenc_sampler_branches_vs_reference_golden(case, golden_dualenc_branches, rxn0, syn4):
    """DualEncoderEpsNetwork.langevin_dynamics_sample with ddpm_noisy (its default), ddpm_det and generalized
    (dualenc.py:861-944) against the reference's own trajectories, same injected noise."""
    g, ref, kw = graph_for(case, rxn0, syn4), golden_dualenc_branches[case], DUALENC_BRANCH_CASES[case]
    m = make_model("dualenc", 0, DEV)
    d = to_dev(g, DEV)
    n_steps = ref["noise"].size(0)
    pos, traj = m.langevin_dynamics_sample(d["atom_type"], ref["pos_init"].to(DEV), d["bond_index"], d["bond_type"],
                                           d["batch"], g["num_graphs"], extend_order=True, n_steps=n_steps,
                                           step_lr=1e-7, noise=ref["noise"], **kw)
    scale = max(1.0, float(ref["traj"].abs().max()))  # the ddpm branches blow positions up at random init
    assert len(traj) == n_steps
    assert (torch.stack(traj) - ref["traj"]).abs().max() < 1e-4 * scale
    assert (pos.cpu() - ref["pos"]).abs().max() < 1e-4 * scale


def test_rigid_motion_invariance_batch100():
    """Size-independent property at BASELINE config-2 size: the score network only sees distances and
    bond orders, so a rigid motion of every reaction (one rotation, per-reaction translations) leaves the
    edge set and edge_inv unchanged (up to fp32 rounding of the rotated coordinates) and rotates the
    per-atom score eq_transform(edge_inv)."""
    g = make_batch(100, seed=4)
    m = make_model("condensenc", 0, DEV)
    d = to_dev(g, DEV)
    torch.manual_seed(21)
    pos = (g["pos_init"] * 4.0).to(DEV)
    q, _ = torch.linalg.qr(torch.randn(3, 3, dtype=torch.float64))
    if torch.det(q) < 0:
        q[:, 0] = -q[:, 0]
    shift = torch.randn(100, 3, dtype=torch.float64)[g["batch"]] * 5.0
    pos2 = (pos.double().cpu() @ q.T + shift).float().to(DEV)
    ei1, idx1, ln1 = m(d["atom_type"], d["r_feat"], d["p_feat"], pos, d["bond_index"], d["bond_type"], d["batch"], None)
    ei2, idx2, ln2 = m(d["atom_type"], d["r_feat"], d["p_feat"], pos2, d["bond_index"], d["bond_type"], d["batch"], None)
    assert torch.equal(idx1, idx2), "a pair sits within rounding distance of the cutoff: pick another seed"
    assert rel_err(ln2, ln1) < 1e-5
    assert rel_err(ei2, ei1) < EPS_TOL and max_rel_err(ei2, ei1) < EPS_TOL
    n = pos.size(0)
    s1 = O.eq_transform(ei1.cpu(), pos.cpu(), idx1.cpu(), ln1.cpu())
    s2 = O.eq_transform(ei2.cpu(), pos2.cpu(), idx2.cpu(), ln2.cpu())
    assert s1.shape == (n, 3)
    assert rel_err(s2.double(), s1.double() @ q.T) < 1e-4


@pytest.mark.parametrize("name", ["syn4", "rxn0"])
def test_get_loss_forward_vs_reference_golden(name, golden_loss, rxn0, syn4):
    """get_loss forward values (the validation loss of train.py) of both networks against the reference's own
    get_loss with the same time steps and noise (path B with gradients: tests/test_gpu_training.py; path A with
    gradients enabled refuses: its backward is not built)."""
    g = graph_for(name, rxn0, syn4)
    d = to_dev(g, DEV)
    ref = golden_loss["b_" + name]
    mb = make_model("condensenc", 0, DEV)
    with torch.no_grad():
        loss = mb.get_loss(d["atom_type"], d["r_feat"], d["p_feat"], ref["pos"].to(DEV), d["bond_index"], d["bond_type"],
                           d["batch"], None, g["num_graphs"], time_step=ref["time_step"].to(DEV),
                           pos_noise=ref["pos_noise"].to(DEV))
        drawn = mb.get_loss(d["atom_type"], d["r_feat"], d["p_feat"], ref["pos"].to(DEV), d["bond_index"],
                            d["bond_type"], d["batch"], None, g["num_graphs"])
    assert loss.shape == ref["loss"].shape and rel_err(loss, ref["loss"]) < 1e-4
    assert drawn.shape == ref["loss"].shape and bool(torch.isfinite(drawn).all())
    ref = golden_loss["a_" + name]
    ma = make_model("dualenc", 0, DEV)
    with pytest.raises(NotImplementedError):
        ma.get_loss(d["atom_type"], ref["pos"].to(DEV), d["bond_index"], d["bond_type"], d["batch"], None, g["num_graphs"])
    with torch.no_grad():
        loss, lg, ll = ma.get_loss(d["atom_type"], ref["pos"].to(DEV), d["bond_index"], d["bond_type"], d["batch"], None,
                                   g["num_graphs"], return_unreduced_loss=True, time_step=ref["time_step"].to(DEV),
                                   pos_noise=ref["pos_noise"].to(DEV))
    assert rel_err(lg, ref["loss_global"]) < 1e-4 and rel_err(ll, ref["loss_local"]) < 1e-4
    assert rel_err(loss, ref["loss"]) < 1e-4
