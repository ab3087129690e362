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
