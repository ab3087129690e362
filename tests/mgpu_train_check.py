"""2-GPU check of the data-parallel training step (run under torchrun, one rank per GPU): the 8-reaction batch is cut in
two unequal shards (3 + 5 reactions), every rank runs get_loss(...).mean().backward() on its shard and
training.allreduce_gradients weights by the atom share; every rank also computes the gradient of the WHOLE batch on its
own.  Both must agree (fp32 re-association only: bound 1e-5 of each tensor's norm), and the ranks must end bit-equal."""
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from helpers import make_model, to_dev  # noqa: E402
from tsdiff_b200.synthetic import make_batch  # noqa: E402
from tsdiff_b200.training import allreduce_gradients  # noqa: E402


def grads_of(model, g, dev, time_step, noise):
    d = to_dev(g, dev)
    for p in model.parameters():
        p.grad = None
    loss = model.get_loss(d["atom_type"], d["r_feat"], d["p_feat"], d["pos_init"] * 1.5, d["bond_index"], d["bond_type"],
                          d["batch"], d["num_nodes_per_graph"], g["num_graphs"], time_step=time_step.to(dev),
                          pos_noise=noise.to(dev))
    loss.mean().backward()
    return loss.size(0)


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    assert world == 2
    g = make_batch(8, seed=11)
    gen = torch.Generator().manual_seed(3)
    t_all = torch.randint(0, 5000, (8,), generator=gen)
    z_all = torch.randn(g["atom_type"].numel(), 3, generator=gen)
    model = make_model("condensenc", 0, dev)
    model.train()
    params = [p for p in model.parameters() if p.requires_grad]
    grads_of(model, g, dev, t_all, z_all)
    full = [p.grad.clone() for p in params]
    # unequal shards: reactions [0, 3) and [3, 8)
    cut = 3
    counts = g["num_nodes_per_graph"]
    n0 = int(counts[:cut].sum())
    keep = list(range(cut)) if rank == 0 else list(range(cut, 8))
    # build the shard by masking atoms / bonds of the chosen reactions
    node_mask = torch.isin(g["batch"], torch.tensor(keep))
    remap = torch.cumsum(node_mask.long(), 0) - 1
    bmask = node_mask[g["bond_index"][0]]
    sub = {"atom_type": g["atom_type"][node_mask], "r_feat": g["r_feat"][node_mask], "p_feat": g["p_feat"][node_mask],
           "pos_init": g["pos_init"][node_mask], "bond_index": remap[g["bond_index"][:, bmask]],
           "bond_type": g["bond_type"][bmask], "batch": g["batch"][node_mask] - (0 if rank == 0 else cut),
           "num_nodes_per_graph": counts[keep], "num_graphs": len(keep)}
    n_local = grads_of(model, sub, dev, t_all[keep], z_all[node_mask])
    assert n_local == (n0 if rank == 0 else g["atom_type"].numel() - n0)
    allreduce_gradients(params, n_local)
    worst = 0.0
    for p, f in zip(params, full):
        err = float((p.grad - f).norm() / f.norm().clamp(min=1e-20))
        worst = max(worst, err)
    flat = torch.cat([p.grad.reshape(-1) for p in params])
    both = [torch.empty_like(flat) for _ in range(world)]
    dist.all_gather(both, flat)
    same = bool(torch.equal(both[0], both[1]))
    print("rank %d: atoms %d of %d, worst relative gradient difference vs the whole-batch gradient %.3e, ranks bit-equal: %s"
          % (rank, n_local, g["atom_type"].numel(), worst, same), flush=True)
    dist.barrier()
    dist.destroy_process_group()
    if worst >= 1e-5 or not same:
        raise SystemExit("data-parallel gradients differ from the single-GPU gradient")


if __name__ == "__main__":
    main()
