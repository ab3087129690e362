"""Experiment: the batch of 100 reactions split into K independent shards, each with its own engine and captured
step graph, replayed round-robin on K streams (reactions never interact, Philox noise is keyed by the global atom
id, so the union of the shards' results equals the single-batch result).  Prints samples/s for K = 1, 2, 3, 4."""
import sys, time, torch
sys.path.insert(0, '.')
import bench
from tsdiff_b200.synthetic import shard_batch

class A: pass
args = A(); args.batch = 100; args.network = 'condensenc'; args.math = 'tf32'; args.ld_steps = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
dev = torch.device('cuda:0')
data = bench.build_inputs(args, 0)
model, cfg = bench.make_models(args, dev)
for K in (1, 2, 3, 4):
    runners, streams = [], []
    for r in range(K):
        shard = shard_batch(data, r, K) if K > 1 else data
        dd = {k: (v.to(dev) if torch.is_tensor(v) else v) for k, v in shard.items()}
        eng, runner = bench.build_runner(args, model, dd, keep_traj=False)
        runner.prepare()
        runners.append(runner); streams.append(torch.cuda.Stream(device=dev))
    torch.cuda.synchronize()
    def run(n):
        for s in streams: s.wait_stream(torch.cuda.current_stream())
        for _ in range(n):
            for r, s in zip(runners, streams):
                with torch.cuda.stream(s):
                    r.graph.replay()
        for s in streams: torch.cuda.current_stream().wait_stream(s)
    run(50)
    torch.cuda.synchronize()
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0.record(); run(args.ld_steps); t1.record(); torch.cuda.synchronize()
    ms = t0.elapsed_time(t1)
    print("shards %d: %.1f us per step of all shards -> %.1f samples/s (5000-step trajectories)" % (K, ms * 1e3 / args.ld_steps, 100 / (ms * 1e-3 / args.ld_steps * 5000)), flush=True)
    del runners
