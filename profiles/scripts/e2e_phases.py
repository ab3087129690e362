"""Where the end-to-end overhead of one dynamic_sampling call goes (host wall clock, synchronised per phase)."""
import sys, time, torch
sys.path.insert(0, '.')
import bench
from tsdiff_b200 import engine as E

class A: pass
args = A(); args.batch = 100; args.network = 'condensenc'; args.math = 'tf32'; args.ld_steps = 5000
dev = torch.device('cuda:0')
data = bench.build_inputs(args, 0)
model, cfg = bench.make_models(args, dev)
pinned = {k: (v.pin_memory() if torch.is_tensor(v) else v) for k, v in data.items()}
def phase(name, t0):
    torch.cuda.synchronize(); t = time.perf_counter(); print("%-34s %8.2f ms" % (name, (t - t0) * 1e3), flush=True); return t
for it in range(2):
    print("---- call", it)
    t = time.perf_counter(); t_all = t
    d = {k: (v.to(dev, non_blocking=True) if torch.is_tensor(v) else v) for k, v in pinned.items()}
    t = phase("H2D inputs", t)
    eng = E.CondensedScoreEngine([model], d["atom_type"], d["r_feat"], d["p_feat"], d["bond_index"], d["bond_type"], d["batch"], math='tf32')
    t = phase("engine (plan, K1, weight views)", t)
    sched, sigmas = E.ld_schedule(model.alphas, args.ld_steps, 1e-7)
    pos = (d["pos_init"] * sigmas[-1].to(dev)).contiguous().clone()
    ch0, ch1 = eng.score_channels(1000)
    runner = E.LangevinRunner(eng, ch0, ch1, sched, pos, seed=2022, keep_traj=True)
    t = phase("schedule + runner buffers", t)
    runner.prepare()
    t = phase("warm-up step + graph capture", t)
    runner.run()
    t = phase("5000 replays", t)
    traj = runner.traj_cpu()
    t = phase("trajectory D2H tail (110 MB streamed during the run)", t)
    lst = list(traj.unbind(0))
    t = phase("unbind", t)
    print("%-34s %8.2f ms" % ("total", (t - t_all) * 1e3))
    del runner, eng
