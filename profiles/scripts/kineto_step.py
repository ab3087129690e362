"""Per-kernel device times of REAL CUDA-graph replays (no cache flush, no serialisation), via
torch.profiler (CUPTI).  Prints a per-kernel summary over `n` replayed Langevin steps."""
import sys, collections, torch
sys.path.insert(0, '.')
from torch.profiler import profile, ProfilerActivity
import bench

class A: pass
args = A(); args.batch = 100; args.network = sys.argv[2] if len(sys.argv) > 2 else 'condensenc'; args.math = sys.argv[1] if len(sys.argv) > 1 else 'tf32'; args.ld_steps = 5000
dev = torch.device('cuda:0')
import os
from tsdiff_b200 import _lib as L
if os.environ.get('NODE_TILE'):
    L.load().tsd_tune_node_tile(int(os.environ['NODE_TILE']))
if os.environ.get('STACK_MODE'):
    L.load().tsd_tune_filter_stack(int(os.environ['STACK_MODE']))
if os.environ.get('STACK_CTAS'):
    L.load().tsd_tune_filter_stack_grid(int(os.environ['STACK_CTAS']))
if os.environ.get('NODE_PDL'):
    L.load().tsd_tune_node_pdl(int(os.environ['NODE_PDL']))
data = bench.build_inputs(args, 0)
model, cfg = bench.make_models(args, dev)
data_dev = {k: (v.to(dev) if torch.is_tensor(v) else v) for k, v in data.items()}
eng, runner = bench.build_runner(args, model, data_dev, keep_traj=False)
runner.prepare()
runner.run(n_steps=1500)   # late-trajectory edge counts
torch.cuda.synchronize()
n = 20
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for _ in range(n):
        runner.graph.replay()
    torch.cuda.synchronize()
agg = collections.OrderedDict()
for ev in prof.events():
    if ev.device_type.name != 'CUDA':
        continue
    a = agg.setdefault(ev.name[:86], [0, 0.0])
    a[0] += 1
    a[1] += ev.device_time_total if hasattr(ev, 'device_time_total') else ev.cuda_time_total
tot = sum(a[1] for a in agg.values())
print("E =", eng.plan.edge_count(), " per-step kernel time %.1f us over %d kernels" % (tot / n, sum(a[0] for a in agg.values()) / n))
for k, a in sorted(agg.items(), key=lambda x: -x[1][1]):
    print("%6.1f/step x%5.1f  avg %7.2f us  %5.1f%%  %s" % (a[1] / n, a[0] / n, a[1] / a[0], 100 * a[1] / tot, k))

# timeline of the last replayed step: start offset, duration, stream of every kernel
evs = [ev for ev in prof.events() if ev.device_type.name == 'CUDA']
evs.sort(key=lambda e: e.time_range.start)
per = len(evs) // n
last = evs[-per:]
t0 = last[0].time_range.start
print("\ntimeline of one step (us from the first kernel's start):  start  dur  end  stream  kernel")
for ev in last:
    st = ev.time_range.start - t0
    du = ev.time_range.end - ev.time_range.start
    name = ev.name.replace('(anonymous namespace)::', '').replace('void ', '')
    print("%8.1f %6.1f %8.1f  s%-3s %s" % (st, du, st + du, getattr(ev, 'device_resource_id', '?'), name[:60]))
