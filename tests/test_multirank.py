"""CPU, world_size 2 over gloo: the reaction-sharding logic of the N>1 path (no collective on
the data path; results gathered at the end).  Each rank scores its shard with the oracle; the
gathered per-reaction outputs must equal the unsharded evaluation."""
import os

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import tsdiff_oracle as O
from tsdiff_b200.config import TRAIN_CONFIG_MODEL
from tsdiff_b200.synthetic import make_batch, shard_batch

from helpers import make_model, oracle_params


def _worker(rank, world, port, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(2)
    g = make_batch(5, seed=11)
    sh = shard_batch(g, rank, world)
    p = oracle_params(make_model("condensenc", 0))
    pos = sh["pos_init"] * 4.0
    ei, idx, _ = O.condensenc_forward(p, TRAIN_CONFIG_MODEL, sh["atom_type"], sh["r_feat"], sh["p_feat"], pos,
                                      sh["bond_index"], sh["bond_type"], sh["batch"])
    node_eq = O.eq_transform(ei, pos, idx, O.edge_lengths(pos, idx).unsqueeze(-1))
    gathered = [None] * world
    dist.all_gather_object(gathered, (sh["graph_range"], sh["atom_offset"], node_eq))
    t = torch.tensor([float(rank + 1)])
    dist.all_reduce(t, op=dist.ReduceOp.MAX)  # the max-over-ranks timing reduction of bench.py
    if rank == 0:
        torch.save({"gathered": gathered, "max": t}, os.path.join(out_dir, "out.pt"))
    dist.barrier()
    dist.destroy_process_group()


def test_sharded_scores_equal_unsharded(tmp_path):
    world = 2
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    res = torch.load(os.path.join(str(tmp_path), "out.pt"), weights_only=False)
    assert float(res["max"]) == world
    parts = sorted(res["gathered"], key=lambda x: x[1])
    assert [p[0] for p in parts] == [(0, 2), (2, 5)]
    g = make_batch(5, seed=11)
    p = oracle_params(make_model("condensenc", 0))
    pos = g["pos_init"] * 4.0
    ei, idx, _ = O.condensenc_forward(p, TRAIN_CONFIG_MODEL, g["atom_type"], g["r_feat"], g["p_feat"], pos,
                                      g["bond_index"], g["bond_type"], g["batch"])
    full = O.eq_transform(ei, pos, idx, O.edge_lengths(pos, idx).unsqueeze(-1))
    merged = torch.cat([p[2] for p in parts])
    assert merged.shape == full.shape
    assert (merged - full).abs().max() < 1e-5 * full.abs().max()


def _member_worker(rank, world, port, out_dir):
    """Ensemble-member-per-rank mode (BASELINE config 3): rank r holds member r and the whole batch; every
    step the per-atom scores eq_transform(edge_inv_r / M) are all-reduced, then every rank applies the same
    update (the host-side contract of LangevinRunner(reduce=...), restated with the oracle's arithmetic)."""
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(2)
    gold = torch.load(os.path.join(os.path.dirname(__file__), "golden", "golden_outputs.pt"), weights_only=False)
    rxn0 = torch.load(os.path.join(os.path.dirname(__file__), "golden", "rxn0_graph.pt"), weights_only=False)
    ref = gold["b_rxn0_ens2_ld5"]
    p = oracle_params(make_model("condensenc", rank))  # members were built with seeds 0 and 1
    sig = O.sigmas_of(p["alphas"])
    T, n_steps = p["alphas"].numel(), ref["noise"].size(0)
    pos = ref["pos_init"] * sig[-1]
    for k, i in enumerate(reversed(range(T - n_steps, T))):
        ei, idx, length = O.condensenc_forward(p, TRAIN_CONFIG_MODEL, rxn0["atom_type"], rxn0["r_feat"], rxn0["p_feat"],
                                               pos, rxn0["bond_index"], rxn0["bond_type"], rxn0["batch"])
        node_eq = O.eq_transform(ei / world, pos, idx, length)
        dist.all_reduce(node_eq)  # the one exchange of the step: (N,3) floats
        pos = O.center_pos(O.ld_update(pos, O.clip_norm(node_eq, 1000), ref["noise"][k], sig[i], 1e-7), rxn0["batch"])
    out = [None] * world
    dist.all_gather_object(out, pos)
    if rank == 0:
        torch.save({"pos": out, "ref": ref["pos"]}, os.path.join(out_dir, "member.pt"))
    dist.barrier()
    dist.destroy_process_group()


def test_member_per_rank_ensemble_matches_reference_golden(tmp_path):
    world = 2
    port = 31500 + (os.getpid() % 2000)
    mp.spawn(_member_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    res = torch.load(os.path.join(str(tmp_path), "member.pt"), weights_only=False)
    assert torch.equal(res["pos"][0], res["pos"][1]), "ranks must stay in lockstep (same sums, same update)"
    assert (res["pos"][0] - res["ref"]).abs().max() < 1e-4  # Angstrom, vs the reference's 2-member ensemble run
