"""BASELINE config 3: N-checkpoint ensemble, one member per GPU, batch 100 -- run under torchrun.
Every rank holds member `rank` and the same batch; the per-atom scores are all-reduced every step inside the
captured step graph.  Prints us per Langevin step and samples/s (100 reactions, 5000-step trajectories)."""
import os, sys, torch
import torch.distributed as dist
sys.path.insert(0, '.')
import bench
from tsdiff_b200 import engine as E
from tsdiff_b200.models.epsnet import get_model
from tsdiff_b200.config import TRAIN_CONFIG_MODEL

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
class A: pass
args = A(); args.batch = 100; args.network = 'condensenc'; args.math = 'tf32'; args.ld_steps = 5000
data = bench.build_inputs(args, 0)                      # the SAME batch on every rank
torch.manual_seed(rank)                                 # a different ensemble member per rank
model = get_model(TRAIN_CONFIG_MODEL).to(dev); model.math = 'tf32'
d = {k: (v.to(dev) if torch.is_tensor(v) else v) for k, v in data.items()}
eng = E.CondensedScoreEngine([model], d["atom_type"], d["r_feat"], d["p_feat"], d["bond_index"], d["bond_type"], d["batch"], math='tf32')
sched, sigmas = E.ld_schedule(model.alphas, args.ld_steps, 1e-7)
ch0, ch1 = eng.score_channels(1000)
pos = (d["pos_init"] * sigmas[-1].to(dev)).contiguous()
runner = E.LangevinRunner(eng, ch0, ch1, sched, pos, seed=2022, keep_traj=False,
                          reduce=lambda t: dist.all_reduce(t), ensemble_size=world)
runner.prepare()
torch.cuda.synchronize(); print('rank', rank, 'captured', flush=True)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1500
for warm in (5, 50, 300):
    runner.run(n_steps=warm)
    torch.cuda.synchronize(); print('rank', rank, 'ran', warm, flush=True)
dist.barrier()
t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
t0.record(); runner.run(n_steps=n); t1.record(); torch.cuda.synchronize()
t = torch.tensor([t0.elapsed_time(t1)], device=dev); dist.all_reduce(t, op=dist.ReduceOp.MAX)
us = float(t.item()) * 1e3 / n
allpos = [torch.empty_like(runner.pos) for _ in range(world)]
dist.all_gather(allpos, runner.pos)
same = all(torch.equal(allpos[0], p) for p in allpos)
if rank == 0:
    print("ensemble of %d members, one per GPU, batch 100: %.1f us per Langevin step (max over ranks) -> %.1f samples/s; ranks in lockstep: %s"
          % (world, us, 100 / (us * 1e-6 * 5000), same), flush=True)
# the captured step graph holds NCCL kernel nodes: release it before the communicator goes away
# (destroy_process_group() with the graph alive hung the teardown)
del runner, eng
torch.cuda.synchronize()
dist.barrier(); dist.destroy_process_group()
