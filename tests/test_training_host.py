"""CPU (gloo, world size 2): the data-parallel gradient reduction of the training step.  The loss is a per-atom mean
(train.py:140-143), so the single-process gradient of the concatenated batch is the ATOM-weighted mean of the
per-rank gradients -- checked with a toy model whose gradients are exact."""
import os

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from tsdiff_b200.training import allreduce_gradients


def _toy_loss(w, x):
    return ((x @ w) ** 2).sum(-1, keepdim=True)  # per-"atom" loss


def _worker(rank, world, port, out_dir):
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.manual_seed(0)
    w = torch.nn.Parameter(torch.randn(5, 3))
    x = torch.randn(11, 5)
    shard = x[:4] if rank == 0 else x[4:]  # 4 and 7 atoms: unequal shares
    _toy_loss(w, shard).mean().backward()
    n = allreduce_gradients([w], shard.size(0))
    assert n == w.numel()
    torch.save(w.grad.clone(), os.path.join(out_dir, "g%d.pt" % rank))
    dist.destroy_process_group()


def test_node_weighted_gradient_allreduce_equals_single_process(tmp_path):
    port = 29650 + os.getpid() % 200
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    torch.manual_seed(0)
    w = torch.nn.Parameter(torch.randn(5, 3))
    x = torch.randn(11, 5)
    _toy_loss(w, x).mean().backward()
    g0, g1 = torch.load(tmp_path / "g0.pt"), torch.load(tmp_path / "g1.pt")
    assert torch.equal(g0, g1)
    assert torch.allclose(g0, w.grad, rtol=1e-6, atol=1e-7)
