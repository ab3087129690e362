"""2-GPU check of the ensemble-member-per-GPU mode (run under torchrun, one rank per GPU; spawned by
tests/test_gpu_kernels.py::test_member_per_gpu_ensemble when two devices are visible):
rank r holds member r (seed r) and the whole rxn_0 batch; EnsembleSampler.dynamic_sampling with
ensemble_group= exchanges the per-atom scores every step -- fused into the update kernel over peer-mapped memory
(default) or as an NCCL all-reduce captured in the step's CUDA graph -- and must reproduce the golden trajectory of
the reference's own 2-member EnsembleSampler run.  Also times both exchanges at batch 100 (profiles/)."""
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from helpers import make_model, to_dev  # noqa: E402
from tsdiff_b200.models.sampler import EnsembleSampler  # noqa: E402


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    gold = torch.load(os.path.join(ROOT, "tests", "golden", "golden_outputs.pt"), weights_only=False)
    rxn0 = torch.load(os.path.join(ROOT, "tests", "golden", "rxn0_graph.pt"), weights_only=False)
    ref = gold["b_rxn0_ens2_ld5"]
    assert world == 2
    d = to_dev(rxn0, dev)
    worst = 0.0
    for exchange in ("fused", "nccl"):
        for use_graph in (False, True):
            ens = EnsembleSampler([make_model("condensenc", rank, dev)])
            pos, traj = ens.dynamic_sampling(d["atom_type"], d["r_feat"], d["p_feat"], ref["pos_init"].to(dev),
                                             d["bond_index"], d["bond_type"], d["batch"], 1, extend_order=True,
                                             n_steps=ref["noise"].size(0), step_lr=1e-7, clip=1000, sampling_type="ld",
                                             noise=ref["noise"], use_graph=use_graph, ensemble_group=dist.group.WORLD,
                                             ensemble_exchange=exchange)
            err = float((torch.stack(traj) - ref["traj"]).abs().max())
            worst = max(worst, err)
            both = [torch.empty_like(pos) for _ in range(world)]
            dist.all_gather(both, pos)
            assert torch.equal(both[0], both[1]), "ranks diverged"
            print("rank %d exchange=%s use_graph=%s max |traj - reference| = %.3e" % (rank, exchange, use_graph, err),
                  flush=True)
    dist.barrier()
    dist.destroy_process_group()
    if worst >= 1e-4:
        raise SystemExit("member-per-GPU ensemble differs from the reference golden trajectory: %.3e" % worst)


if __name__ == "__main__":
    main()
